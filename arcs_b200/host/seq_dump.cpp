// seq_dump.cpp -- test helper: prints what SeqReader (seq_reader.h) parses from a file, one record per
// line as  <ret>\t<name>\t<comment>\t<seq>\t<qual-length>, then the terminating return value.
// Used by tests/test_host_cpu.py to pin the host parser against an independent restatement of the
// record grammar (Arcs/kseq.h:175-215).
#include "seq_reader.h"
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv)
{
	if (argc < 2)
		return 2;
	arks_host::SeqReader rd(argv[1], argc > 2 ? (size_t)atoi(argv[2]) : (1u << 20));
	if (!rd.ok())
		return 3;
	arks_host::SeqRecord r;
	int l;
	while ((l = rd.read(r)) >= 0) {
		r.truncate_at_nul();
		printf("%d\t%s\t%s\t%s\t%zu\n", l, r.name.c_str(), r.comment.c_str(), r.seq.c_str(), r.qual.size());
	}
	printf("END %d\n", l);
	return 0;
}
