#!/bin/bash
# Round-2 closing session: test-suite, the default bench line (every BASELINE config + latency), the reference arm, the kernel list.
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -rfEx --timeout 300 > $OUT/r2_final_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/r2_final_pytest.log
tail -4 $OUT/r2_final_pytest.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > $OUT/r2_final_bench_reference.json 2> $OUT/r2_final_bench_reference.err; echo "reference arm rc=$?"
timeout 1500 python bench.py > $OUT/r2_final_bench.json 2> $OUT/r2_final_bench.err; echo "bench rc=$?"; tail -c 300 $OUT/r2_final_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final_bench.json"))
def show(x, name):
    r = x["roofline"]; o = x["roofline_other"]
    print(name, "value", x["value"], "ms", x["ms_per_step"], "enc", x["encode_ms"], "dec", x["decode_ms"], r["kernel"][:8], r["kernel_ms"], o["kernel"][:8], o["kernel_ms"], "e2e", (x.get("e2e") or {}).get("value"), "cpu", (x.get("cpu_baseline") or {}).get("value"), "parity", x.get("parity",{}).get("mismatches"), x.get("parity",{}).get("frames_checked"))
show(d, "c2")
for c in d.get("configs", []): show(c, c["name"])
print(json.dumps(d.get("latency"))[:1200])
r = json.load(open("gpurun_out/r2_final_bench_reference.json"))
print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
bash profiles/r2_kernlist.sh r2_final2 c2
