#!/bin/bash
# Source-level ncu capture of one amp_conv_tc launch (stage 2, C=40, k=3, conv2 with residual) at the
# full bench batch; the .ncu-rep stays on the box (too large for gpurun_out), only CSV extracts return.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BA="--steps 1 --warmup 1 --no-cpu-baseline --no-roofline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:amp_conv_tc_kernel -s 37 -c 1 -o /tmp/ac -f python bench.py $BA > gpurun_out/ncu_acsrc.log 2>&1; echo "exit $?"
ncu -i /tmp/ac.ncu-rep --page raw --csv > gpurun_out/ac_raw.csv 2>/dev/null
ncu -i /tmp/ac.ncu-rep --page source --csv --print-source sass > /tmp/ac_sass.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/ac_sass.csv')))
hdr=rows[1]; body=rows[2:]
ci=hdr.index('Instructions Executed'); si=hdr.index('# Samples')
def val(r,i):
    try: return float(r[i].replace(',',''))
    except: return 0.0
tot=sum(val(r,ci) for r in body); tots=sum(val(r,si) for r in body)
with open('gpurun_out/ac_sass_hot.csv','w') as f:
    w=csv.writer(f); w.writerow(['idx','sass','inst_executed','pct_inst','samples','pct_samples'])
    for i,r in enumerate(body):
        v=val(r,ci); sm=val(r,si)
        if v/tot > 0.002 or sm/max(tots,1) > 0.004: w.writerow([i, r[1].strip(), int(v), round(100*v/tot,2), int(sm), round(100*sm/max(tots,1),2)])
print('total warp-instr', tot, 'samples', tots, 'rows', len(body))
PY
ls -la gpurun_out/
