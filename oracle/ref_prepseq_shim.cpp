// oracle/ref_prepseq_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" window onto the reference's unmodified ReadsProcessor::prepSeq
// (Common/ReadsProcessor.cpp:376-535) so tests can pin oracle/arks_oracle.c's
// canonical-key restatement against it window by window.
#include "Common/ReadsProcessor.h"
#include <cstring>
#include <string>
extern "C" {
// returns key length in bytes, or -1 if prepSeq returned NULL; writes key bytes to out
int
ref_prepseq(const char* seq, int seqlen, int pos, int k, unsigned char* out)
{
	ReadsProcessor proc(k);
	std::string s(seq, seqlen);
	const unsigned char* r = proc.prepSeq(s, pos);
	if (!r)
		return -1;
	std::string key = proc.getStr(r);
	memcpy(out, key.data(), key.size());
	return (int)key.size();
}
// all windows of one sequence: out is n_windows * nb bytes, valid[i] = 0/1
int
ref_prepseq_all(const char* seq, int seqlen, int k, unsigned char* out, unsigned char* valid)
{
	ReadsProcessor proc(k);
	std::string s(seq, seqlen);
	int nb = (k + 3) / 4;
	for (int i = 0; i + k <= seqlen; ++i) {
		const unsigned char* r = proc.prepSeq(s, i);
		valid[i] = r != NULL;
		if (r)
			memcpy(out + (size_t)i * nb, r, nb);
		else
			memset(out + (size_t)i * nb, 0, nb);
	}
	return nb;
}
}
