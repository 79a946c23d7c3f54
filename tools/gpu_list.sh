#!/bin/bash
# Per-kernel time, instruction count and DRAM traffic of two steady-state scans (ncu, few metrics).  Usage: bash tools/gpu_list.sh <tag> [bench args]
tag=${1:-r02x}; shift
out=gpurun_out; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct \
   --clock-control none -s 70 -c 34 --csv --log-file $out/${tag}_list.csv \
   python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-extra "$@" > $out/${tag}_list.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$out/${tag}_list.csv")) if len(r)>10]
hdr=rows[0]; 
ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki].split("(")[0][-40:]),{})[r[mi]]=float(r[vi].replace(",",""))
for (i,k),m in sorted(d.items()):
    print("%3d %-42s %8.1f us  %7.2f Minst  R %7.1f W %7.1f MB  issue %4.1f%% warps %4.1f%% L2hit %4.1f%%"%(i,k,m.get("gpu__time_duration.sum",0)/1e3,m.get("smsp__inst_executed.sum",0)/1e6,m.get("dram__bytes_read.sum",0)/1e6,m.get("dram__bytes_write.sum",0)/1e6,m.get("smsp__issue_active.avg.pct_of_peak_sustained_active",0),m.get("sm__warps_active.avg.pct_of_peak_sustained_active",0),m.get("lts__t_sector_hit_rate.pct",0)))
PY
