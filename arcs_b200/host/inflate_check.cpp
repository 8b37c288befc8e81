// inflate_check.cpp -- test helper: decompresses a .gz with FastInflate (fast_inflate.h) in reads of a given
// size and writes the bytes to stdout; prints the error (if any) and the byte count to stderr.  With "zlib"
// as the first argument it does the same through zlib's gzread, for comparison and timing.
#include "bgzf_inflate.h"
#include "par_inflate.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

int main(int argc, char** argv)
{
	if (argc < 3)
		return 2;
	const bool use_zlib = std::string(argv[1]) == "zlib";
	const bool parallel = std::string(argv[1]).compare(0, 3, "par") == 0; // par<threads>, e.g. par4
	const int par_threads = parallel && argv[1][3] ? atoi(argv[1] + 3) : 4;
	const size_t par_chunk = getenv("PAR_CHUNK") ? (size_t)atol(getenv("PAR_CHUNK")) : (1u << 20);
	const size_t chunk = argc > 3 ? (size_t)atol(argv[3]) : (1u << 20);
	const bool quiet = argc > 4;
	std::vector<char> buf(chunk);
	size_t total = 0;
	auto t0 = std::chrono::steady_clock::now();
	if (use_zlib) {
		gzFile f = gzopen(argv[2], "r");
		if (!f)
			return 3;
		gzbuffer(f, 1u << 20);
		int n;
		while ((n = gzread(f, buf.data(), (unsigned)chunk)) > 0) {
			if (!quiet)
				fwrite(buf.data(), 1, (size_t)n, stdout);
			total += (size_t)n;
		}
		int err;
		const char* msg = gzerror(f, &err);
		if (err != Z_OK && err != Z_STREAM_END)
			fprintf(stderr, "ERROR %s\n", msg);
		gzclose(f);
	} else {
		int fd = open(argv[2], O_RDONLY);
		if (fd < 0)
			return 3;
		struct stat st;
		fstat(fd, &st);
		const uint8_t* m = st.st_size ? (const uint8_t*)mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
		if (std::string(argv[1]).compare(0, 4, "bgzf") == 0) { // bgzf<threads>
			const bool is = arks_host::BgzfInflate::is_bgzf(m, (size_t)st.st_size);
			fprintf(stderr, "IS_BGZF %d\n", (int)is);
			arks_host::BgzfInflate inf(m, (size_t)st.st_size, argv[1][4] ? atoi(argv[1] + 4) : 4, par_chunk);
			long n;
			while ((n = inf.read(buf.data(), chunk)) > 0) {
				if (!quiet)
					fwrite(buf.data(), 1, (size_t)n, stdout);
				total += (size_t)n;
			}
			if (!inf.ok())
				fprintf(stderr, "ERROR %s\n", inf.error().c_str());
			fprintf(stderr, "PARALLEL_MEMBERS %zu\n", inf.parallel_members());
		} else if (parallel) {
			arks_host::ParInflate inf(m, (size_t)st.st_size, par_threads, par_chunk);
			long n;
			while ((n = inf.read(buf.data(), chunk)) > 0) {
				if (!quiet)
					fwrite(buf.data(), 1, (size_t)n, stdout);
				total += (size_t)n;
			}
			if (!inf.ok())
				fprintf(stderr, "ERROR %s\n", inf.error().c_str());
			fprintf(stderr, "PARALLEL_CHUNKS %zu\n", inf.parallel_chunks());
		} else {
			arks_host::FastInflate inf(m, (size_t)st.st_size);
			long n;
			while ((n = inf.read(buf.data(), chunk)) > 0) {
				if (!quiet)
					fwrite(buf.data(), 1, (size_t)n, stdout);
				total += (size_t)n;
			}
			if (!inf.ok())
				fprintf(stderr, "ERROR %s\n", inf.error().c_str());
		}
	}
	const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	fprintf(stderr, "BYTES %zu  %.3f s  %.1f MB/s\n", total, s, total / 1e6 / s);
	return 0;
}
