set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r1_final2_pytest_gpu.txt
timeout 900 python bench.py > gpurun_out/r1_final2_bench_1gpu.json 2> gpurun_out/r1_final2_bench_1gpu.err
LGTEUN_TIMING=1 timeout 300 python bench.py --no-graph --batch 256 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline 2> gpurun_out/timing_final2.txt > /dev/null
tail -19 gpurun_out/timing_final2.txt > gpurun_out/r1_final2_event_timing_gf2_b256.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_final2_launches_gf2_b16.csv python bench.py --batch 16 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"ffn_tc_kernel|window_msa_kernel" --launch-skip 4 -c 2 -o gpurun_out/r1_final2_top2 python bench.py --no-graph --batch 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --skip-dead-priors > gpurun_out/ncu_final2.log 2>&1
cat gpurun_out/r1_final2_pytest_gpu.txt
cut -c1-330 gpurun_out/r1_final2_bench_1gpu.json
