#!/usr/bin/env bash
# Build oracle/_ref/arcs_ref: the reference's OWN hot-path code (bcgsc/arcs 1.2.8),
# compiled from the sources where they lie under /root/reference.
#
# TEST INFRASTRUCTURE ONLY.  Nothing in the product path may execute this binary.
#
# The full reference cannot be built here (Boost, google-sparsehash, btllib and
# autotools are absent, no network).  Its hot-path functions do compile verbatim:
# this recipe pulls the line ranges listed below out of Arcs/Arcs.{h,cpp} into a
# throw-away translation unit in a temp dir (never into the repo), compiles them
# together with the unmodified Common/ReadsProcessor.cpp and Arcs/kseq.h, and
# keeps only the resulting binary in oracle/_ref/ (git-ignored, gpurun-shipped).
# Shims (in ref_driver.cpp, our code):  google::sparse_hash_map -> std::unordered_map
# (the map is never iterated by the reference, so the container is not
# result-bearing) and a Boost-free createGraph/write_graphviz restatement.
set -euo pipefail
REF="${ARCS_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/Arcs/Arcs.cpp" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT (if any)" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT

# anchor checks: refuse to build if the reference moved under us
chk() { # file line pattern
  sed -n "${2}p" "$REF/$1" | grep -q -- "$3" || { echo "build_ref: anchor mismatch $1:$2 ($3)" >&2; exit 1; }
}
chk Arcs/Arcs.h 29 'namespace ARCS'
chk Arcs/Arcs.h 128 'ContigToLengthIt'
chk Arcs/Arcs.cpp 179 's_numkmersmapped'
chk Arcs/Arcs.cpp 187 'KSEQ_INIT'
chk Arcs/Arcs.cpp 237 'calcJacIndex'
chk Arcs/Arcs.cpp 318 'checkContigSequence'
chk Arcs/Arcs.cpp 367 'checkReadSequence'
chk Arcs/Arcs.cpp 393 'createIndexMultMap'
chk Arcs/Arcs.cpp 834 'normalEstimation'
chk Arcs/Arcs.cpp 870 'mapKmers'
chk Arcs/Arcs.cpp 940 'bestContig'
chk Arcs/Arcs.cpp 1014 '^}'
chk Arcs/Arcs.cpp 1022 'getContigKmers'
chk Arcs/Arcs.cpp 1133 'chromiumRead'
chk Arcs/Arcs.cpp 1379 'pairContigs'
chk Arcs/Arcs.cpp 1460 'checkSignificance'
chk Arcs/Arcs.cpp 1710 'writeTSV'
chk Arcs/Arcs.cpp 206 'checkFlag'
chk Arcs/Arcs.cpp 280 'calcSequenceIdentity'
chk Arcs/Arcs.cpp 573 'readBAM'
chk Arcs/Arcs.cpp 771 '^}'
chk Arcs/Arcs.cpp 800 'readBAMS'
chk Arcs/DistanceEst.h 16 'struct DistanceEstimate'
chk Arcs/DistanceEst.h 102 'calcDistSamples'
chk Arcs/DistanceEst.h 338 'estimateDistance'
chk Arcs/DistanceEst.h 389 '^}'
chk Arcs/DistanceEst.h 502 'writeDistSamplesTSV'

x() { sed -n "${2},${3}p" "$REF/$1"; }
{ x Arcs/Arcs.h 29 128; } > "$TMP/ref_types.inc"          # namespace ARCS { ... (driver closes it)
{ x Arcs/Arcs.cpp 179 185; x Arcs/Arcs.cpp 187 202; x Arcs/Arcs.cpp 235 254;
  x Arcs/Arcs.cpp 316 331; x Arcs/Arcs.cpp 363 389; x Arcs/Arcs.cpp 391 547;
  x Arcs/Arcs.cpp 814 830; x Arcs/Arcs.cpp 832 1014; } > "$TMP/ref_part_a.inc"
{ x Arcs/Arcs.cpp 1015 1370; x Arcs/Arcs.cpp 1372 1467; x Arcs/Arcs.cpp 1674 1757; } > "$TMP/ref_part_b.inc"
# -D distance estimation: everything of Arcs/DistanceEst.h that does not touch the Boost graph
# (types, calcDistSamples, buildJaccardToDist, buildPairToBarcodeStats, estimateDistance, the
# samples writer); addEdgeDistances / writeDistTSV are restated over the driver's own graph
{ x Arcs/DistanceEst.h 15 389; x Arcs/DistanceEst.h 495 535; } > "$TMP/ref_dist.inc"
# ARCS alignment mode: the SAM record loop that fills the IndexMap (checkFlag, checkChar,
# calcSequenceIdentity, readBAM, readBAMS); getScaffSizes (DataLayer/FastaReader needs the autotools
# config.h) is restated in the driver over kseq
{ x Arcs/Arcs.cpp 204 220; x Arcs/Arcs.cpp 275 315; x Arcs/Arcs.cpp 572 771; x Arcs/Arcs.cpp 799 812; } > "$TMP/ref_sam.inc"

g++ -std=c++11 -O2 -fopenmp -w -I"$TMP" -I"$REF" -I"$REF/Common" -I"$REF/Arcs" \
    "$HERE/ref_driver.cpp" "$REF/Common/ReadsProcessor.cpp" -lz -o "$OUT/arcs_ref"
# the key canonicaliser alone, as a shared object, for unit-level pinning of the restatement
g++ -std=c++11 -O2 -w -shared -fPIC -I"$REF" -I"$REF/Common" \
    "$HERE/ref_prepseq_shim.cpp" "$REF/Common/ReadsProcessor.cpp" -o "$OUT/libref_prepseq.so"
echo "build_ref: built $OUT/arcs_ref $OUT/libref_prepseq.so"
