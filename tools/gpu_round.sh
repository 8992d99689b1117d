#!/bin/bash
# tools/gpu_round.sh <tag> -- the standard sequence of a GPU session (run under gpurun from the repository root):
# gpu tests, tile-kernel A/B, bench config 3 (the metric) and the other BASELINE configs, ncu launch list.
# Everything lands in gpurun_out/<tag>_*.
T=${1:-rXX}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; tail -6 $O/${T}_pytest.log
timeout 300 python tools/tile_ab.py --cfgs ${TILE_CFGS:-0,5} --quick --base 0 > $O/${T}_tile_ab.log 2>&1; grep "^cfg\|mismatch\|identical" $O/${T}_tile_ab.log | cut -c1-220
timeout 500 python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err; tail -2 $O/${T}_bench_n1.err
python - <<PY
import json
try:
    d=json.load(open("$O/${T}_bench_n1.json"))
    print("N1", round(d["value"]), round(d["e2e"]["value"]), "ms/step %.3f" % d["ms_per_step"], "tile %.3f" % d["roofline"]["kernel_ms"], "frac %.5f" % d["roofline"]["frac"])
    print("  sustained", d["sustained"], "ceiling", d["e2e"]["h2d_ceiling"]["frames_per_s"], "clocks", d["clocks"])
    print("  parity", d.get("parity")); print("  cpu", d.get("cpu_baseline")); print("  stages", d["stage_ms_per_batch"], "numa", d["config"]["numa"])
except Exception as ex: print("bench n1 failed", ex)
PY
if [ "${SKIP_CONFIGS:-0}" != "1" ]; then
for c in 1 2 4 5; do
  timeout 500 python bench.py --config $c > $O/${T}_bench_c$c.json 2> $O/${T}_bench_c$c.err; tail -2 $O/${T}_bench_c$c.err
  python - <<PY
import json
try:
    d=json.load(open("$O/${T}_bench_c$c.json"))
    print("C$c", d["metric"], "value %.1f e2e %.1f ms/step %.3f frac %.5f cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("cpu_baseline",{}).get("value")))
    if "svm" in d: print("   svm", d["svm"])
    if "levels" in d: print("   levels", [(l["width"], l["planes"], round(l["tile_ms"],3), l["kept_nodes"]) for l in d["levels"]], "latency", d["e2e"].get("latency_ms_one_frame_alone"))
except Exception as ex: print("bench c$c failed", ex)
PY
done
fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 3 --contexts 1 --no-cpu-baseline --no-next-rows --sustained-seconds 0 --jpeg-seconds 0 --parity-frames 0 > $O/${T}_ncu_b.log 2>&1
python - <<PY
import csv, collections, re, statistics
try:
    lines=[l for l in open("$O/${T}_launches.csv") if not l.startswith("==")]
    r=csv.reader(lines); hdr=next(r); ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
    tot=collections.defaultdict(list)
    for row in r:
        if len(row)<=vi: continue
        v=float(row[vi].replace(",","")); v = v/1000 if row[ui]=="ns" else (v*1000 if row[ui]=="ms" else v)
        tot[re.sub(r"\(.*","",row[ki])].append(v)
    s=sum(statistics.median(v) for v in tot.values())
    for k,v in sorted(tot.items(), key=lambda kv:-statistics.median(kv[1])): print("  %-36s %8.1f us %5.1f%%" % (k[:36], statistics.median(v), 100*statistics.median(v)/s))
    print("  sum %.1f us" % s)
except Exception as ex: print("launch list failed", ex)
PY
