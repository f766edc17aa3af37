#!/bin/bash
# shared-memory wavefronts per LDS of the FFMA2 inner-loop probes (per-instruction, from the source page)
set -u
mkdir -p gpurun_out
for m in ${NCU_MODES:-4 8 0 16 20 24}; do
  timeout 200 ncu --set full --clock-control none --import-source on -k "regex:probe_ffma2_lds" -s 3 -c 1 -o gpurun_out/prof_probe_$m -f python tools/probe_ffma2.py $m > gpurun_out/ncu_probe_$m.log 2>&1
  ncu -i gpurun_out/prof_probe_$m.ncu-rep --page source --csv 2>/dev/null | python -c "
import csv,sys,collections
rows=list(csv.reader(sys.stdin)); hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
agg=collections.defaultdict(lambda:[0,0])
for r in rows[2:]:
    src=r[ix['Source']].split()
    if not src: continue
    op=src[0] if not src[0].startswith('@') else src[1]
    if op.startswith('LDS'):
        agg[op][0]+=int(r[ix['Instructions Executed']] or 0); agg[op][1]+=int(r[ix['L1 Wavefronts Shared']] or 0)
print('mode $m', {k:(v[0],v[1],round(v[1]/max(v[0],1),2)) for k,v in agg.items()})
"
  rm -f gpurun_out/prof_probe_$m.ncu-rep
done
timeout 200 python tools/probe_ffma2.py ${RUN_MODES:-0 16 18 20 24 22}
