for bn in 0 128 192 256; do for st in 0 2 3; do
  echo "== BN=$bn STAGES=$st"
  MSU_TC_BN=$bn MSU_TC_STAGES=$st python tools/gemm_case.py 20 plain_s2 fc1_s2 dxn_s2 2>&1 | grep -E "_s2"
done; done
