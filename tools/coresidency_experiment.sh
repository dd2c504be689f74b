for f in "" "-DMPCB_EVAL_MAXNREG=224 -DMPCB_TGT_BLOCK=32" "-DMPCB_EVAL_MAXNREG=224 -DMPCB_TGT_BLOCK=32 -DKKT_WARPS=2"; do
  export MPCB_EXTRA_FLAGS="$f"
  echo "flags: [$f]"
  MPCB_BENCH_NOSAMPLER=1 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('  bench joined G=2: value %.0f ms %.3f shares %s' % (d['value'], d['ms_per_step'], {k: round(v,3) for k,v in d['roofline']['kernel_time_share'].items() if v>0.05}))"
  for g in 2 3; do python tools/two_streams.py $g 100 2>&1 | tail -1 | sed 's/^/  free-running: /'; done
done
