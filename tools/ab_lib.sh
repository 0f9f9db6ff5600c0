L=semantic_segmentation_of_stylegan2_artifacts_b200/libmsunet_sm100.so
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo NEW; timeout 200 python tools/gemm_case.py 10 2>&1 | grep " us "; timeout 100 python tools/wgrad_case.py 10 | grep wgrad; timeout 100 python tools/attn_case.py 10 0 1 | grep stage
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
cp $L /tmp/new.so; cp tools/build/lib_old.so $L
echo OLD; timeout 200 python tools/gemm_case.py 10 2>&1 | grep " us "; timeout 100 python tools/wgrad_case.py 10 | grep wgrad; timeout 100 python tools/attn_case.py 10 0 1 | grep stage
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
cp /tmp/new.so $L
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
