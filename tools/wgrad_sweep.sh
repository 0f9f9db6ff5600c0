for mt in 1 2 4; do for bn in 128 256; do
echo "MT=$mt BN=$bn"
MSU_WG_MT=$mt MSU_WG_BN=$bn python tools/wgrad_case.py 10 16384,1536,384 19600,1152,384 16384,384,1536 4096,3072,768 7056,2304,768 65536,768,192 19600,384,384 4096,768,768 2>&1 | grep wgrad
done; done
