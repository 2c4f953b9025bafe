#!/bin/bash
# Round-2 measurement pass on one B200 box (run through gpurun from the repo root):
#   gpurun --timeout 2400 -- 'bash profiles/r2_capture.sh <tag>'
# Everything lands in gpurun_out/<tag>_*; profiles/summarise.py turns the ncu output into the
# committed summaries.  ncu passes run only after the same command has exited 0 without ncu.
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
NCU_BENCH="python bench.py --steps 1 --warmup 1 --batches-per-step 2 --pool 2 --no-cpu-baseline --no-other-configs --skew-heavy 0"

nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_gputest.log 2>&1
echo "pytest rc=$?" | tee -a $out/${tag}_gputest.log
tail -3 $out/${tag}_gputest.log

timeout 1200 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
echo "reference arm rc=$?"

# parity sweeps (CUDA path vs oracle, every check of tests/test_gpu_parity.py)
: > $out/${tag}_parity_sweep.jsonl
timeout 600 python tests/gpu_parity_sweep.py 240 300 16 >> $out/${tag}_parity_sweep.jsonl 2>> $out/${tag}_sweep.err
timeout 600 python tests/gpu_parity_sweep.py 64 24 16 large >> $out/${tag}_parity_sweep.jsonl 2>> $out/${tag}_sweep.err
timeout 600 python tests/gpu_parity_sweep.py 240 48 16 csr >> $out/${tag}_parity_sweep.jsonl 2>> $out/${tag}_sweep.err
timeout 600 python tests/gpu_parity_sweep.py 120 90 16 motor >> $out/${tag}_parity_sweep.jsonl 2>> $out/${tag}_sweep.err
echo "sweeps done"

# files -> frames (host decoding, device-side decoding)
timeout 400 python tests/gpu_pipeline_bench.py > $out/${tag}_pipeline.jsonl 2> $out/${tag}_pipeline.err
echo "pipeline rc=$?"

# launch list + full captures of the top kernels
if timeout 600 $NCU_BENCH > $out/${tag}_ncu_plain.json 2> $out/${tag}_ncu_plain.err; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
      --log-file $out/${tag}_launches.csv $NCU_BENCH > $out/${tag}_ncu1.log 2>&1
  echo "launch list rc=$?"
  timeout 1200 ncu --set full --clock-control none --import-source on \
      -k regex:'gmm_em_kernel|gmm_init_kernel|gmm_final_kernel|position_kernel_warp|ks_small_kernel|ks_wave_kernel' \
      -s 28 -c 10 -f -o $out/${tag}_top $NCU_BENCH > $out/${tag}_ncu2.log 2>&1
  echo "full capture rc=$?"
fi
ls -la $out
