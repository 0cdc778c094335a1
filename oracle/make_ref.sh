#!/bin/sh
# Stages the UNMODIFIED hot-path modules of the reference next to the zuko / sbi stand-ins, so that
# `bench.py --impl reference` (and the cpu_baseline leg) can time the reference's own `NSE.ddim` on the GPU box,
# where /root/reference does not exist.  Outputs go to oracle/_ref/ only: git-ignored (reference sources are never
# committed), not gpurun-ignored (it travels like the built .so).  TEST / MEASUREMENT INFRASTRUCTURE ONLY -- the
# product path never imports it (tests/test_cabi_cpu.py::test_product_never_imports_the_oracle).
#   usage: sh oracle/make_ref.sh [reference checkout, default /root/reference]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-${D4S_REFERENCE:-/root/reference}}"
OUT="$HERE/_ref"
if [ ! -f "$REF/nse.py" ]; then
    echo "make_ref: no reference checkout at $REF (nothing staged)" >&2
    exit 0
fi
rm -rf "$OUT"
mkdir -p "$OUT"
# the six files of the sampling path
for f in nse.py pf_nse.py tall_posterior_sampler.py vp_diffused_priors.py nse_toy.py embedding_nets.py; do
    cp "$REF/$f" "$OUT/$f"
done
cp -r "$HERE/ref_shims/zuko" "$OUT/zuko"
cp -r "$HERE/ref_shims/sbi" "$OUT/sbi"
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo unknown ) > "$OUT/REFERENCE_COMMIT"
( cd "$OUT" && sha256sum nse.py pf_nse.py tall_posterior_sampler.py vp_diffused_priors.py nse_toy.py embedding_nets.py ) > "$OUT/SHA256SUMS"
echo "make_ref: staged $(ls "$OUT" | wc -l) entries in $OUT"
