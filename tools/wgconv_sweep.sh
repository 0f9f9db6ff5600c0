for cfg in "0 0" "1 0" "1 3" "1 2" "0 2"; do
  set -- $cfg
  echo "== HALO=$1 STAGES=$2"
  MSU_WGRAD_HALO=$1 MSU_WGRAD_STAGES=$2 python tools/wgrad_conv_case.py 2>&1 | grep wgrad
done
