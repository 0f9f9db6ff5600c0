set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1.err
python bench.py --steps 3 --warmup 3 --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1000 --csv --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:gemm_tc_kernel -c 1 -o gpurun_out/r02_conv3x3_tc python tools/gemm_case.py 1 conv_b16 > gpurun_out/ncu_conv.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:winattn -c 2 -o gpurun_out/r02_winattn_tc python tools/attn_case.py 1 0 > gpurun_out/ncu_attn.log 2>&1
python tools/attn_case.py 10 > gpurun_out/r02_attn_cases.txt 2>&1
python tools/ln_case.py > gpurun_out/r02_ln_cases.txt 2>&1
python tools/gemm_case.py 10 > gpurun_out/r02_gemm_cases.txt 2>&1
python tools/wgrad_case.py 10 > gpurun_out/r02_wgrad_cases.txt 2>&1
python tools/wgrad_conv_case.py > gpurun_out/r02_wgrad_conv_cases.txt 2>&1
python tools/microbench.py > gpurun_out/r02_microbench.jsonl 2> gpurun_out/r02_microbench.err
python tools/timeline.py gpurun_out/r02_timeline.json > /dev/null 2>&1
python tools/op_table.py > gpurun_out/r02_op_table.txt 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --img 1024 --batch 8 > gpurun_out/r02_bench_n1_b8_1024.json 2>/dev/null
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --attn-drop 0.05 > gpurun_out/r02_bench_n1_attn_drop_0.05.json 2>/dev/null
