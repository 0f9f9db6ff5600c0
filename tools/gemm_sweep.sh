# stage-2 GEMM tiling / CTA-pair sweep (tools/gemm_case.py cases)
C="fc1g_s2 dmul_s2 plain_s2 fc2_s2 dxn_s2 qkv_s2 proj_s2 fc1_s1 fc2_s1"
echo default; python tools/gemm_case.py 10 $C 2>&1 | grep " us "
echo PAIR_K=384; MSU_TC_PAIR_K=384 python tools/gemm_case.py 10 $C 2>&1 | grep " us "
echo PAIR_K=192; MSU_TC_PAIR_K=192 python tools/gemm_case.py 10 $C 2>&1 | grep " us "
for bn in 128 256; do echo BN=$bn; MSU_TC_BN=$bn python tools/gemm_case.py 10 $C 2>&1 | grep " us "; done
echo BN=256 PAIR_K=384; MSU_TC_BN=256 MSU_TC_PAIR_K=384 python tools/gemm_case.py 10 $C 2>&1 | grep " us "
