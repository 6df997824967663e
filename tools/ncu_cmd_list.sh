#!/bin/bash
# ncu launch list of an arbitrary command.  Usage: bash tools/ncu_cmd_list.sh TAG cmd...
TAG=$1; shift; O=gpurun_out; mkdir -p $O
"$@" > $O/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_$TAG.csv "$@" > $O/ncu_list_$TAG.log 2>&1
echo "ncu list rc=$?"; tail -2 $O/plain_$TAG.log | cut -c1-300
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$O/launches_$TAG.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    u=r[ui]; v = v/1e3 if u=="ns" else (v if u in ("us","usecond") else v*1e3 if u=="ms" else v/1e3)
    k=r[ki].split("(")[0]; a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:25]: print(f"{t/c:10.1f} us x {c:4d}  {k}")
PY
