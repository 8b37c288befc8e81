// seq_dump.cpp -- test helper: prints what SeqReader (seq_reader.h) parses from a file, one record per
// line as  <ret>\t<name>\t<comment>\t<seq>\t<qual-length>, then the terminating return value.
// Used by tests/test_host_cpu.py to pin the host parser against an independent restatement of the
// record grammar (Arcs/kseq.h:175-215).
#include "fasta_fast.h"
#include "seq_reader.h"
#include <cstdio>
#include <cstdlib>
#include <string>
int main(int argc, char** argv)
{
	if (argc < 2)
		return 2;
	// seq_dump --fast <file> <threads>: the records as the all-cores FASTA path (fasta_fast.h) lists them -- name, sequence,
	// and the first / last min(7, len / 2) bases through copy_bases -- or FALLBACK if the file is not in the strict shape
	if (argc > 2 && std::string(argv[1]) == "--fast") {
		arks_host::MappedFasta mf;
		if (!mf.open(argv[2], argc > 3 ? atoi(argv[3]) : 4)) {
			printf("FALLBACK\n");
			return 0;
		}
		for (const auto& r : mf.records()) {
			std::string seq(r.seq_len, '?');
			arks_host::MappedFasta::copy_bases(r, 0, r.seq_len, &seq[0]);
			const size_t cut = std::min<size_t>(7, r.seq_len / 2);
			std::string head(cut, '?'), tail(cut, '?');
			arks_host::MappedFasta::copy_bases(r, 0, cut, &head[0]);
			arks_host::MappedFasta::copy_bases(r, r.seq_len - cut, cut, &tail[0]);
			printf("%zu\t%s\t%s\t%s\t%s\n", r.seq_len, std::string(r.name, r.name_n).c_str(), seq.c_str(), head.c_str(), tail.c_str());
		}
		printf("END -1\n");
		return 0;
	}
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
