#!/bin/bash
# GPU-box visit 2: full parity tests, bench line, SF10 parity vs the reference module, ring (fixed) vs LDG A/B, ncu
set -u
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "smoke failed"; exit 1; fi
echo "== tests" ; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" ; tail -8 gpurun_out/tests.log
echo "== bench" ; timeout 600 python bench.py > gpurun_out/bench_q1.json 2> gpurun_out/bench_q1.err; echo "bench rc=$?"; cat gpurun_out/bench_q1.json
echo "== A/B ring vs LDG (SF10)"
timeout 900 python tools/ab_variants.py --sf 10 --queries q1,q6,q3,q5,q9,q18 --variants default,ring --out gpurun_out/ab2.json > gpurun_out/ab2.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/ab2.log
echo "== SF10 check vs reference"
timeout 1200 python tools/run_tpch.py --sf 10 --queries q1,q6,q3,q5,q9,q18 --check --out gpurun_out/sf10_check.json > gpurun_out/sf10_check.log 2>&1; echo "rc=$?"; cut -c1-330 gpurun_out/sf10_check.log
echo "== pcie"
timeout 120 python - <<'PY' > gpurun_out/pcie.txt 2>&1
import torch, time
h = torch.empty(660_000_000, dtype=torch.uint8, pin_memory=True)
d = torch.empty_like(h, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print("pinned H2D 660 MB: %.2f ms  %.1f GB/s" % (dt * 1e3, 0.66 / dt))
PY
cat gpurun_out/pcie.txt
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/q1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full q1_k0 / q6_k0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q1_k0 -s 3 -c 1 -o gpurun_out/q1_k0 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:q6_k0 -s 3 -c 1 -o gpurun_out/q6_k0 -f python bench.py --query q6 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full6.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -40
