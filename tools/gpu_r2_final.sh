#!/bin/bash
# round 2, final evidence run on one B200: tests, the bench lines, launch list, ncu --set full on C2 / C3 / C4, the reference's protocol
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv > gpurun_out/f_smi.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log; tail -4 gpurun_out/f_pytest.log
timeout 900 python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_n1_reference.json 2> gpurun_out/f_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --workload c2_cornell --steps 5 > gpurun_out/f_bench_c2.json 2> gpurun_out/f_bench_c2.err; echo "c2 rc=$?"
timeout 900 python bench.py --workload c4_heightfield_10m --steps 5 > gpurun_out/f_bench_c4.json 2> gpurun_out/f_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
for f in ("f_bench_n1", "f_bench_c2", "f_bench_c4", "f_bench_n1_reference"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, {k: round(v["mrays_per_s"], 1) for k, v in d.get("renderers", {}).items()},
              "e2e", d.get("e2e") and round(d["e2e"]["value"], 1), "frac", d.get("roofline") and round(d["roofline"]["frac"], 3), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 2))
    except Exception as e:
        print(f, "FAILED", e)
PY
# launch list of the bench command (one metric, no replay)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches_bench.csv python bench.py --no-cpu-baseline > gpurun_out/f_ncu_launches.log 2>&1; echo "launch list rc=$?"
prof() { # workload renderer spp kernel-regex tag
  python tools/profile_run.py --workload $1 --renderer $2 --spp $3 --frames 2 > gpurun_out/f_plain_$5.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$4 -s 1 -c 1 -o gpurun_out/r02c_$5 -f python tools/profile_run.py --workload $1 --renderer $2 --spp $3 --frames 2 > gpurun_out/f_ncu_$5.log 2>&1
  tail -2 gpurun_out/f_plain_$5.log | head -1
}
prof c3_sponza_scale megakernel 32 k_megakernel mega_c3
prof c3_sponza_scale wavefront 32 k_wf_flow flow_c3
prof c2_cornell megakernel 64 k_megakernel mega_c2
prof c2_cornell wavefront 64 k_wf_flow flow_c2
prof c4_heightfield_10m megakernel 16 k_megakernel mega_c4
prof c4_heightfield_10m wavefront 16 k_wf_flow flow_c4
timeout 1500 python tools/benchmark_protocol.py gpurun_out/r02_benchmark_protocol.csv 3 > gpurun_out/f_protocol.log 2>&1; tail -3 gpurun_out/f_protocol.log
