// tools/micro/pack_bench.cpp -- how fast can the host turn ASCII bases into the 2-bit words + invalid
// mask the lookup kernel works on?  (Decides whether packing before the H2D copy can beat PCIe.)
//   g++ -O3 -fopenmp -o pack_bench pack_bench.cpp && ./pack_bench [GB] [threads]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>
#include <omp.h>
#include <vector>

// 32 bases -> two 32-bit words (16 bases each, base 0 in bits 31-30) + 32-bit invalid mask
__attribute__((target("avx2"))) static inline void pack32(const char* src, uint32_t* w, uint32_t* inv, uint32_t* n_n)
{
	const __m256i x = _mm256_loadu_si256((const __m256i*)src);
	const __m256i m3 = _mm256_set1_epi8(3), m1 = _mm256_set1_epi8(1);
	__m256i c = _mm256_xor_si256(_mm256_and_si256(_mm256_srli_epi16(x, 1), m3), _mm256_and_si256(_mm256_srli_epi16(x, 2), m1));
	const __m256i y = _mm256_or_si256(x, _mm256_set1_epi8(0x20));
	const __m256i lut = _mm256_setr_epi8('a', 'c', 'g', 't', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'a', 'c', 'g', 't', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
	const __m256i ok = _mm256_cmpeq_epi8(y, _mm256_shuffle_epi8(lut, c));
	*inv = ~(uint32_t)_mm256_movemask_epi8(ok);
	*n_n = (uint32_t)__builtin_popcount((uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(y, _mm256_set1_epi8('n'))));
	c = _mm256_and_si256(c, ok);
	// 4 codes -> one byte, first base in the top bits
	const __m256i p16 = _mm256_maddubs_epi16(c, _mm256_set1_epi32(0x01041040)); // bytes 64,16,4,1
	const __m256i p32 = _mm256_madd_epi16(p16, _mm256_set1_epi16(1));            // one packed byte per 32-bit lane
	// lanes 0-3 / 4-7 -> bytes; base 0 of each word must land in the most significant byte
	const __m256i sh = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
	const __m256i b = _mm256_shuffle_epi8(p32, sh);
	w[0] = (uint32_t)_mm256_extract_epi32(b, 0);
	w[1] = (uint32_t)_mm256_extract_epi32(b, 4);
}

int main(int argc, char** argv)
{
	const double gb = argc > 1 ? atof(argv[1]) : 2.0;
	const int threads = argc > 2 ? atoi(argv[2]) : omp_get_max_threads();
	const size_t read_len = 150, n_reads = (size_t)(gb * 1e9 / read_len);
	std::vector<char> bases(n_reads * read_len + 64);
#pragma omp parallel for num_threads(threads)
	for (size_t i = 0; i < bases.size(); ++i)
		bases[i] = "ACGT"[(i * 2654435761u >> 7) & 3];
	const size_t words_per_read = (read_len + 15) / 16;
	std::vector<uint32_t> W(n_reads * words_per_read + 8), INV(n_reads * ((read_len + 31) / 32) + 8), NB(n_reads);
	for (int rep = 0; rep < 3; ++rep) {
		auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for num_threads(threads) schedule(static)
		for (size_t r = 0; r < n_reads; ++r) {
			const char* s = bases.data() + r * read_len;
			uint32_t* w = W.data() + r * words_per_read;
			uint32_t* iv = INV.data() + r * ((read_len + 31) / 32);
			uint32_t nn = 0;
			size_t i = 0;
			for (; i + 32 <= read_len; i += 32) {
				uint32_t n;
				pack32(s + i, w + i / 16, iv + i / 32, &n);
				nn += n;
			}
			if (i < read_len) { // tail through a padded copy
				char tmp[32];
				memset(tmp, 'A', 32);
				memcpy(tmp, s + i, read_len - i);
				uint32_t ww[2], n;
				pack32(tmp, ww, iv + i / 32, &n);
				w[i / 16] = ww[0];
				if (i / 16 + 1 < words_per_read)
					w[i / 16 + 1] = ww[1];
				nn += n;
			}
			NB[r] = nn;
		}
		double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		printf("threads=%d  %.2f GB of bases packed in %.3f s = %.1f GB/s\n", threads, n_reads * read_len / 1e9, s, n_reads * read_len / 1e9 / s);
	}
	unsigned long long chk = 0;
	for (size_t i = 0; i < W.size(); i += 9973)
		chk += W[i] + INV[i % INV.size()];
	printf("checksum %llu\n", chk);
	return 0;
}
