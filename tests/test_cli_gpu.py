"""The `arcs --arks` drop-in command line (arcs_b200/bin/arcs) against
 - the reference's golden demo outputs (Examples/arks_test-demo, arks-long_test-demo) and
 - outputs of the reference's own code (oracle/_ref/arcs_ref) on adversarial inputs, committed under
   tests/golden/cli_cases by tools/make_fixtures.py.
Byte-for-byte on _original.gv, _main.tsv, barcode counts and the pair map."""
import glob
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
ARCS = os.path.join(ROOT, "arcs_b200", "bin", "arcs")


def run_arcs(args, cwd):
    p = subprocess.run([ARCS] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p.stdout


def read(path):
    with open(path) as f:
        return f.read()


def test_arks_demo_cli(tmp_path):
    d = os.path.join(GOLD, "arks_demo")
    out = run_arcs(["--arks", "-v", "-f", os.path.join(d, "test_scaffolds.renamed.fa"), "-c", "5", "-m", "50-6000", "-r", "0.05",
                    "-e", "30000", "-z", "500", "-j", "0.55", "-k", "30", "-t", "8", "-d", "0", "--gap", "100", "-b",
                    str(tmp_path / "demo"), os.path.join(d, "test_reads.fq.gz"), "--barcode-counts", str(tmp_path / "bc.tsv")],
                   tmp_path)
    assert read(tmp_path / "demo_original.gv") == read(os.path.join(d, "expected_original.gv"))
    # verbose counters of the golden log
    for line in ["Total number of Kmers:  123190", "Number Null Kmers:  303", "Number Kmers Recorded:  118710",
                 "Number Kmer Collisions:  4480", "Number Times Kmers Removed (since duplicate in different contig):  547",
                 "Number of unique kmers (only one contig):  118334", "Stored read pairs: 21632",
                 "Skipped reads pairs without a good contig: 6012", "Total valid kmers: 6109324",
                 "Number of kmers found in ContigKmap: 4862376", "Number of kmers recorded in Ktrack: 4814099",
                 "Number of kmers found in ContigKmap but duplicate: 48277", "Number of reads passing jaccard threshold: 44503",
                 "Number of reads failing jaccard threshold: 10785"]:
        assert line in out, line
    assert os.path.exists(tmp_path / "demo.dist.gv") and os.path.exists(tmp_path / "demo_main.tsv")
    assert len(read(tmp_path / "bc.tsv").splitlines()) == 1085


def test_arks_long_demo_cli(tmp_path):
    d = os.path.join(GOLD, "arks_long_demo")
    run_arcs(["--arks", "-v", "-f", os.path.join(d, "test_scaffolds.renamed.fa"), "-c", "3", "-m", "8-10000", "-r", "0.05", "-e",
              "30000", "-z", "500", "-j", "0.05", "-k", "20", "-t", "8", "-d", "0", "--gap", "100", "-b", str(tmp_path / "long"),
              "-u", os.path.join(d, "barcodeMultiplicityArcs.tsv"), os.path.join(d, "test_reads.cut250.fq.gz")], tmp_path)
    assert read(tmp_path / "long_original.gv") == read(os.path.join(d, "expected_original.gv"))
    assert read(tmp_path / "long_main.tsv") == read(os.path.join(d, "expected_main.tsv"))
    # .dist.gv lists vertices in the iteration order of a std::unordered_map<std::string,int>
    # (Arcs.cpp:1622), which depends on the libstdc++ build: with this image's g++ 13 the reference
    # itself gives "2 3 1", the golden (older toolchain) "3 2 1".  Same lines, order aside:
    got = read(tmp_path / "long.dist.gv").splitlines()
    want = read(os.path.join(d, "expected.dist.gv")).splitlines()
    assert got[0] == want[0] == "digraph arcs {" and got[-1] == want[-1] == "}"
    assert sorted(got) == sorted(want)
    nv = sum(1 for x in want if " -> " not in x and x.startswith('"'))
    assert all(" -> " not in x for x in got[1:1 + nv]) and all(" -> " in x for x in got[1 + nv:-1])
    # ... and byte for byte what the reference's own code writes on this toolchain (tools/make_dist_gv_fixtures.py)
    assert read(tmp_path / "long.dist.gv") == read(os.path.join(d, "expected_refcode.dist.gv"))


def test_arks_long_demo_stdin(tmp_path):
    """arcs-make's arks-long rule pipes reads into /dev/stdin with -u (bin/arcs-make:302-313)"""
    d = os.path.join(GOLD, "arks_long_demo")
    cmd = ("zcat %s | %s --arks -f %s -c 3 -m 8-10000 -r 0.05 -e 30000 -z 500 -j 0.05 -k 20 -b %s -u %s /dev/stdin"
           % (os.path.join(d, "test_reads.cut250.fq.gz"), ARCS, os.path.join(d, "test_scaffolds.renamed.fa"), tmp_path / "s",
              os.path.join(d, "barcodeMultiplicityArcs.tsv")))
    subprocess.run(cmd, shell=True, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=300)
    assert read(tmp_path / "s_original.gv") == read(os.path.join(d, "expected_original.gv"))


# (the long-read cases of cut_k20 have their own file, test_zz_cut_gpu.py)
CASES = sorted(c for c in glob.glob(os.path.join(GOLD, "cli_cases", "*", "expected_*_args.json")) if "cut" not in json.load(open(c)))


@pytest.mark.parametrize("argsfile", CASES, ids=[os.path.relpath(c, os.path.join(GOLD, "cli_cases")) for c in CASES])
@pytest.mark.parametrize("mode", ["one-pass", "two-pass", "2gpu-ids", "3gpu-ids"])
def test_cli_matches_reference_code(argsfile, mode, tmp_path):
    d = os.path.dirname(argsfile)
    tag = os.path.basename(argsfile)[len("expected_"):-len("_args.json")]
    spec = json.load(open(argsfile))
    exp = os.path.join(d, "expected_" + tag)
    arcs_mode = spec.get("mode") == "arcs"  # alignment mode (SAM text in, Arcs.cpp:572-771)
    if arcs_mode and mode != "one-pass":
        pytest.skip("read-ingest modes do not apply to alignment input")
    args = ["-b", str(tmp_path / "o"), "--barcode-counts", str(tmp_path / "bc.tsv"), "-P"] + spec["args"]
    if not arcs_mode:
        args = ["--arks", "-f", os.path.join(d, "draft.fa")] + args
    elif spec.get("with_f"):
        args = ["-f", os.path.join(d, "draft.fa")] + args
    if spec["multfile"]:
        args += ["-u", os.path.join(d, spec["multfile"])]
    dist = "-D" in spec["args"]
    if dist:  # distance estimation (Arcs/DistanceEst.h): per-edge estimates and the intra-contig samples
        args += ["--dist_tsv", str(tmp_path / "dist.tsv"), "--samples_tsv", str(tmp_path / "samples.tsv")]
    if mode == "two-pass":
        args.append("--two-pass")
    env_gpus = None
    if mode in ("2gpu-ids", "3gpu-ids"):
        # exercises the barcode-sharded multi-handle path (sharded blocks, one host thread per shard, the merge of
        # the pair-link maps); with fewer physical GPUs the shards land on device 0
        env_gpus = mode[0]
    args.append(os.path.join(d, "aln.sam" if arcs_mode else "reads.fq.gz"))
    env = dict(os.environ)
    if env_gpus:
        env["ARKS_GPUS"] = env_gpus
        import torch
        if torch.cuda.device_count() < int(env_gpus):  # on a multi-GPU box the shards go to real devices (NCCL merge)
            env["ARKS_GPUS_SAME_DEVICE"] = "1"
    p = subprocess.run([ARCS] + args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert read(tmp_path / "o_original.gv") == read(exp + "_original.gv")
    assert read(tmp_path / "o_main.tsv") == read(exp + "_main.tsv")
    assert read(tmp_path / "o_pair.tsv") == read(exp + "_pmap.txt")
    assert read(tmp_path / "bc.tsv") == read(exp + "_bc.tsv")
    # vertex order = the walk over the reference's unordered_map contigToLength (Arcs.cpp:1622), pinned by its own code
    assert read(tmp_path / "o.dist.gv") == read(exp + "_dist.gv")
    if dist:
        assert read(tmp_path / "dist.tsv") == read(exp + "_dist.tsv")
        assert read(tmp_path / "samples.tsv") == read(exp + "_samples.tsv")
        # the ABySS graph carries the estimate (or INT_MAX where there is none) instead of the fixed gap
        est = {ln.split("\t")[3] for ln in read(exp + "_dist.tsv").splitlines()[1:]}
        want_d = {x for x in est if x != "NA"} | ({"2147483647"} if "NA" in est or " weight=" in "".join(
            ln for ln in read(exp + "_original.gv").splitlines() if "--" in ln and "d=" not in ln) else set())
        got_d = {ln.split("[d=")[1].split(" ")[0] for ln in read(tmp_path / "o.dist.gv").splitlines() if " -> " in ln}
        assert got_d == want_d


@pytest.mark.parametrize("how", ["two-files", "fof", "sequential-reader"])
def test_cli_several_read_files(how, tmp_path):
    """the reads split over a plain and a gzip file (at a pair boundary) give what the reference's code gives on
    the single file (checked with oracle/_ref when the fixture was made); also through -a, and with the
    block-parallel parser switched off"""
    import gzip
    d = os.path.join(GOLD, "cli_cases", "mixed_k30")
    exp = os.path.join(d, "expected_a")
    spec = json.load(open(exp + "_args.json"))
    lines = gzip.open(os.path.join(d, "reads.fq.gz"), "rb").read().split(b"\n")
    h = (len(lines) // 2) // 8 * 8
    (tmp_path / "a.fq").write_bytes(b"\n".join(lines[:h]) + b"\n")
    with gzip.open(tmp_path / "b.fq.gz", "wb") as f:
        f.write(b"\n".join(lines[h:]))
    args = ["--arks", "-f", os.path.join(d, "draft.fa"), "-b", str(tmp_path / "o"), "--barcode-counts", str(tmp_path / "bc.tsv"),
            "-P"] + spec["args"]
    env = dict(os.environ)
    if how == "fof":
        (tmp_path / "reads.fof").write_text("%s\n%s\n" % (tmp_path / "a.fq", tmp_path / "b.fq.gz"))
        args += ["-a", str(tmp_path / "reads.fof")]
    else:
        args += [str(tmp_path / "a.fq"), str(tmp_path / "b.fq.gz")]
    if how == "sequential-reader":
        env["ARKS_PARSE_THREADS"] = "0"
    p = subprocess.run([ARCS] + args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert read(tmp_path / "o_original.gv") == read(exp + "_original.gv")
    assert read(tmp_path / "o_main.tsv") == read(exp + "_main.tsv")
    assert read(tmp_path / "o_pair.tsv") == read(exp + "_pmap.txt")
    assert read(tmp_path / "bc.tsv") == read(exp + "_bc.tsv")


@pytest.mark.parametrize("n_gpus", [2, 4, 8])
def test_cli_output_is_invariant_in_the_number_of_gpus(n_gpus, tmp_path):
    """integer sums => `.gv`, `_main.tsv`, the pair map and the barcode counts are byte-identical at
    1/2/4/8 GPUs.  Real devices, NCCL merge; skipped when the box has fewer."""
    import torch
    if torch.cuda.device_count() < n_gpus:
        pytest.skip("needs %d GPUs" % n_gpus)
    d = os.path.join(GOLD, "cli_cases", "plain_k60")
    spec = json.load(open(sorted(glob.glob(os.path.join(d, "expected_*_args.json")))[0]))
    outs = {}
    for n in (1, n_gpus):
        args = ["--arks", "-v", "-f", os.path.join(d, "draft.fa"), "-b", str(tmp_path / ("o%d" % n)), "--barcode-counts",
                str(tmp_path / ("bc%d.tsv" % n)), "-P", "--gpus", str(n)] + spec["args"] + [os.path.join(d, "reads.fq.gz")]
        p = subprocess.run([ARCS] + args, cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
        outs[n] = [read(tmp_path / ("o%d%s" % (n, x))) for x in ("_original.gv", "_main.tsv", "_pair.tsv", ".dist.gv")] + [
            read(tmp_path / ("bc%d.tsv" % n))]
    assert outs[1] == outs[n_gpus]
    assert len(outs[1][2].splitlines()) > 0
