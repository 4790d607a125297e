set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r1_final3_pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_final3_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/r1_final3_bench_1gpu.json 2> gpurun_out/r1_final3_bench_1gpu.err
LGTEUN_TIMING=1 timeout 300 python bench.py --no-graph --batch 256 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline 2> gpurun_out/timing_final3.txt > /dev/null
tail -19 gpurun_out/timing_final3.txt > gpurun_out/r1_final3_event_timing_gf2_b256.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_final3_launches_gf2_b16.csv python bench.py --batch 16 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"fre_cols_fused_kernel|fft_rows_fwd_pre_kernel|fft_rows_inv_post_kernel" --launch-skip 9 -c 3 -o gpurun_out/r1_final3_freprocess python bench.py --workload freprocess --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final3_fre.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"win_attn" --launch-skip 3 -c 1 -o gpurun_out/r1_final3_winattn python bench.py --workload winattn --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_final3_win.log 2>&1
cat gpurun_out/r1_final3_pytest_gpu.txt
tail -2 gpurun_out/r1_final3_smoke.txt
cut -c1-330 gpurun_out/r1_final3_bench_1gpu.json
