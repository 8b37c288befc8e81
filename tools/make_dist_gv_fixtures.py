#!/usr/bin/env python
"""expected_<tag>_dist.gv for every fixture under tests/golden/cli_cases (and the arks-long demo), written by the
reference's own code (oracle/_ref/arcs_ref --dist-gv): the vertex order of `.dist.gv` is the iteration order of the
reference's std::unordered_map<std::string,int> contigToLength (Arcs/Arcs.cpp:1622), which the extracted code
reproduces on this image's libstdc++.  Run here (needs oracle/_ref, i.e. /root/reference once)."""
import glob
import json
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "arcs_ref")


def main():
    for af in sorted(glob.glob(os.path.join(ROOT, "tests/golden/cli_cases/*/expected_*_args.json"))):
        spec = json.load(open(af))
        d = os.path.dirname(af)
        base = os.path.join(d, "expected_" + os.path.basename(af)[len("expected_"):-len("_args.json")])
        if "cut" in spec:
            continue  # the cut reads are made at test time; their .dist.gv is covered by the set comparison
        with tempfile.TemporaryDirectory() as tmp:
            if spec.get("mode") == "arcs":
                cmd = [REF, "--arcs", "-b", os.path.join(tmp, "o")] + spec["args"]
                if spec.get("with_f"):
                    cmd += ["-f", os.path.join(d, "draft.fa")]
                inp = os.path.join(d, "aln.sam")
            else:
                cmd = [REF, "-f", os.path.join(d, "draft.fa"), "-b", os.path.join(tmp, "o")] + spec["args"]
                inp = os.path.join(d, "reads.fq.gz")
            if spec["multfile"]:
                cmd += ["-u", os.path.join(d, spec["multfile"])]
            if "-D" in spec["args"]:
                cmd += ["--dist_tsv", os.path.join(tmp, "d.tsv"), "--samples_tsv", os.path.join(tmp, "s.tsv")]
            subprocess.check_call(cmd + ["--dist-gv", base + "_dist.gv", inp], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            assert open(os.path.join(tmp, "o_original.gv")).read() == open(base + "_original.gv").read(), af
    d = os.path.join(ROOT, "tests/golden/arks_long_demo")
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call([REF, "-f", os.path.join(d, "test_scaffolds.renamed.fa"), "-c", "3", "-m", "8-10000", "-r", "0.05", "-e",
                               "30000", "-z", "500", "-j", "0.05", "-k", "20", "-t", "8", "-d", "0", "--gap", "100", "-b",
                               os.path.join(tmp, "o"), "-u", os.path.join(d, "barcodeMultiplicityArcs.tsv"), "--dist-gv",
                               os.path.join(d, "expected_refcode.dist.gv"), os.path.join(d, "test_reads.cut250.fq.gz")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert open(os.path.join(tmp, "o_original.gv")).read() == open(os.path.join(d, "expected_original.gv")).read()


if __name__ == "__main__":
    main()
