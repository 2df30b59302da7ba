#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BA="--batch 4 --steps 1 --warmup 1 --no-cpu-baseline --no-roofline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp_block_fused -s 9 -c 1 -o /tmp/ab -f python bench.py $BA > gpurun_out/ncu_absrc.log 2>&1; echo "exit $?"
ncu -i /tmp/ab.ncu-rep --page source --csv --print-source sass > /tmp/ab_sass.csv 2>/dev/null
ncu -i /tmp/ab.ncu-rep --page source --csv > /tmp/ab_src.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/ab_sass.csv')))
hdr=rows[1]; body=rows[2:]
ci=hdr.index('Instructions Executed'); si=hdr.index('# Samples')
def val(r,i):
    try: return float(r[i].replace(',',''))
    except: return 0.0
tot=sum(val(r,ci) for r in body); tots=sum(val(r,si) for r in body)
# contiguous hot regions: print every instruction with >0.25% of executed instructions, in address order
with open('gpurun_out/ab_sass_hot.csv','w') as f:
    w=csv.writer(f); w.writerow(['idx','sass','inst_executed','pct_inst','samples','pct_samples'])
    for i,r in enumerate(body):
        v=val(r,ci)
        if v/tot > 0.0025: w.writerow([i, r[1].strip(), int(v), round(100*v/tot,2), int(val(r,si)), round(100*val(r,si)/max(tots,1),2)])
print('total warp-instr', tot, 'rows', len(body))
PY
ls -la gpurun_out/ab_sass_hot.csv
