#!/bin/bash
# whole-codec path: the two full-model GPU tests, then bench config 6 (short) for the stage / family breakdown
mkdir -p gpurun_out; cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_transforms.py -q -m gpu -s -k "full_reference or codec_object" > gpurun_out/r2_codec_tests.log 2>&1; echo "codec tests rc=$?"; grep -E "passed|failed|error" gpurun_out/r2_codec_tests.log | tail -n 3
grep -E "accelerated stack|free-running|codec:|Error|error|assert" gpurun_out/r2_codec_tests.log | head -n 40
timeout 900 python bench.py --config 6 --steps 5 --warmup 3 > gpurun_out/r2_bench_c6.json 2> gpurun_out/r2_bench_c6.err; echo "bench c6 rc=$?"; tail -n 5 gpurun_out/r2_bench_c6.err
python - <<'P'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_c6.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step", "e2e", "gpu_launches_per_step", "stages_ms", "kernel_families", "roofline", "torch_gpu_baseline", "cpu_baseline", "clocks"):
        print(k, json.dumps(d.get(k))[:400])
except Exception as e:
    print("no line", e)
P
