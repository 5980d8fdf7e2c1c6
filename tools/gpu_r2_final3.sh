#!/bin/bash
# round 2, final build (drain carry-over, 72-register megakernel, three rolled node steps per vote): tests, bench lines, reference arm, ncu capture of the megakernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/z_pytest.log; tail -3 gpurun_out/z_pytest.log
timeout 900 python bench.py > gpurun_out/z_bench_n1.json 2> gpurun_out/z_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/z_bench_n1_reference.json 2> gpurun_out/z_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --workload c2_cornell --steps 5 --no-cpu-baseline > gpurun_out/z_bench_c2.json 2> gpurun_out/z_bench_c2.err; echo "c2 rc=$?"
timeout 900 python bench.py --workload c4_heightfield_10m --steps 5 --no-cpu-baseline > gpurun_out/z_bench_c4.json 2> gpurun_out/z_bench_c4.err; echo "c4 rc=$?"
python - <<'PY'
import json
for f in ("z_bench_n1", "z_bench_c2", "z_bench_c4", "z_bench_n1_reference"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, {k: round(v["mrays_per_s"], 1) for k, v in d.get("renderers", {}).items()},
              "e2e", d.get("e2e") and round(d["e2e"]["value"], 1), "frac", d.get("roofline") and round(d["roofline"]["frac"], 3), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 2))
    except Exception as e:
        print(f, "FAILED", e)
PY
prof() { # workload renderer spp kernel-regex tag
  python tools/profile_run.py --workload $1 --renderer $2 --spp $3 --frames 2 > gpurun_out/z_plain_$5.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -s 1 -c 1 -o gpurun_out/r02g_$5 -f python tools/profile_run.py --workload $1 --renderer $2 --spp $3 --frames 2 > gpurun_out/z_ncu_$5.log 2>&1
  tail -2 gpurun_out/z_plain_$5.log | head -1
}
prof c3_sponza_scale megakernel 32 k_megakernel mega_c3
prof c3_sponza_scale wavefront 32 k_wf_flow flow_c3
