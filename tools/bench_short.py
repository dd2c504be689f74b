import json,sys
d=json.loads(sys.stdin.read())
r=d["roofline"]
print(sys.argv[1], "slow", [(a, round(b,1)) for a,b in d.get("slowest_steps", [])], "value %.0f e2e %.0f p50 %.3f frac %.3f avg_eval_ms %.4f shares %s" % (d["value"], d["e2e"]["value"], d["p50_step_latency_ms"], r["frac"], r["avg_launch_ms"], {k: round(v,3) for k,v in r["kernel_time_share"].items()}))
