# which of its data movements bounds the pair conv?  (timing only: MSU_CONV_SKIP makes the results wrong)
# bit 1: no weight loads, bit 2: no halo-row loads, bit 4: no output stores (after each CTA's first tile)
for s in ${1:-0 1 2 3 4 7}; do echo "MSU_CONV_SKIP=$s"; MSU_CONV_SKIP=$s timeout 60 python tools/gemm_case.py 5 conv_b16 2>&1 | grep " us "; done
