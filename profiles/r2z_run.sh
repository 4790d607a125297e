set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2z_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.txt 2>&1
timeout 900 python bench.py > gpurun_out/r2z_bench_1gpu.json 2> gpurun_out/r2z_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference_arm.json 2> /dev/null
tools/opbench --batch 64 > gpurun_out/r2z_opbench_gf2_b64.txt
tools/opbench --bands 8 --batch 32 > gpurun_out/r2z_opbench_wv3_b32.txt
LGTEUN_TIMING=1 tools/opbench --batch 64 --ops forward_nograph --iters 1 > /dev/null 2> gpurun_out/r2z_event_timing_gf2_b64.txt
LGTEUN_TIMING=1 tools/opbench --bands 8 --batch 32 --ops forward_nograph --iters 1 > /dev/null 2> gpurun_out/r2z_event_timing_wv3_b32.txt
python tools/train_prof.py --pan 256 > gpurun_out/r2z_train_kernels_pan256_b4.txt 2> /dev/null
python tools/train_prof.py --pan 128 > gpurun_out/r2z_train_kernels_pan128_b4.txt 2> /dev/null
timeout 300 python bench.py --workload train256 --steps 10 --warmup 3 > gpurun_out/r2z_bench_train256.json 2> /dev/null
timeout 300 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r2z_bench_train128.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches_gf2_b16.csv python bench.py --batch 16 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-eager --no-train-leg --no-parity > /dev/null 2>&1
cat gpurun_out/r2z_pytest_gpu.txt
tail -2 gpurun_out/r2z_smoke.txt
cut -c1-300 gpurun_out/r2z_bench_1gpu.json
head -3 gpurun_out/r2z_train_kernels_pan256_b4.txt
cut -c1-200 gpurun_out/r2z_bench_train256.json
