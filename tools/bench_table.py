#!/usr/bin/env python
import json, sys
d = json.load(open(sys.argv[1]))
print(f"value {d['value']:.0f} {d['unit']}  ms/step {d['ms_per_step']:.3f}  e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.3f} ms, sync call {d['e2e'].get('sync_call_ms', 0):.3f} ms)  launches {d['gpu_launches']}  cpu {d['cpu_baseline']}")
print("roofline", d["roofline"])
for k, v in d["kernels"].items():
    print(f"  {k:24s} {v['us']:8.1f} us  {v['share']:.3f}  {v['tflops']:.1f} TF/s")
