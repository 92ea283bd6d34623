#!/bin/bash
# Rebuild libsdeb200.so and NVRTC-compile (no GPU needed) every kernel the GPU tests and the bench use into
# stochastic_b200/_cubin_cache/, which travels to the GPU box with the repo snapshot.
set -e
cd "$(dirname "$0")"
rm -rf stochastic_b200/_cubin_cache
python __graft_entry__.py build | tail -1
out=$(SDEB200_PRECOMPILE_ONLY=1 python -m pytest tests -m gpu -q -n 8 -p no:cacheprovider 2>&1 || true)
if echo "$out" | grep -q "NVRTC"; then echo "$out" | grep -B2 -A8 "NVRTC" | head -40; echo "NVRTC COMPILE ERRORS"; exit 1; fi
echo "$out" | tail -1
ls stochastic_b200/_cubin_cache/*.cubin | wc -l
