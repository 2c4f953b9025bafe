#!/bin/bash
# Launch list and full capture of the device-side decoder (csrc/decode.cu) on one B200:
#   gpurun --timeout 900 -- 'bash profiles/r2_capture_decode.sh r2f'
# The command is the pipeline bench on a short task list (8 files' worth of transcripts, 64 per batch);
# ncu passes run only after the same command has exited 0 without ncu.
tag=${1:-r2f}
out=gpurun_out
mkdir -p $out
CMD="python tests/gpu_pipeline_bench.py --transcripts 8 --repeat 24 --device-batches 64 --device-only"
if timeout 300 $CMD > $out/${tag}_decode_plain.jsonl 2> $out/${tag}_decode_plain.err; then
  cat $out/${tag}_decode_plain.jsonl
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $out/${tag}_decode_launches.csv $CMD > $out/${tag}_decode_ncu1.log 2>&1
  echo "launch list rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on \
      -k regex:'inflate_kernel|read_kernel|fill_kernel|walk_kernel|expand_runs_kernel' -s 12 -c 8 -f \
      -o $out/${tag}_decode $CMD > $out/${tag}_decode_ncu2.log 2>&1
  echo "full capture rc=$?"
else
  echo "plain run failed"; tail -5 $out/${tag}_decode_plain.err
fi
ls -la $out | tail -8
