import os, sys, subprocess
# A/B of environment knobs, each in a fresh process (tools/perf_probe.py, first line only)
for env in [{}, {"XS_NO_TILED": "1"}, {}, {"XS_NO_TILED": "1"}]:
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "tools/perf_probe.py", "1007000", "70", "only_first"], env=e, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("gemm default")]
    print(env, line[0] if line else out[-300:], flush=True)
for env in [{}, {"XS_NO_TILED": "1"}]:
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "tools/perf_probe.py", "1007000", "70", "big"], env=e, capture_output=True, text=True).stdout
    print(env, [l for l in out.splitlines() if l.startswith("gemm nq")], flush=True)
