#!/bin/bash
# usage: gpu_prof_src2.sh "<kernel regex>:<skip>:<tag>" ...   (bench batch = default 32)
# Source-level ncu captures; reports stay on the box, CSV extracts (key metrics, opcode histogram,
# hot SASS rows) come back under gpurun_out/.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BA="--steps 1 --warmup 1 --no-cpu-baseline --no-roofline"
for spec in "$@"; do
  IFS=: read -r rx skip tag <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o /tmp/$tag -f python bench.py $BA > gpurun_out/ncu_$tag.log 2>&1; echo "$tag exit $?"
  ncu -i /tmp/$tag.ncu-rep --page raw --csv > /tmp/${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/$tag.ncu-rep --page source --csv --print-source sass > /tmp/${tag}_sass.csv 2>/dev/null
  TAG=$tag python - <<'PY'
import csv, os, collections
tag=os.environ['TAG']
rows=list(csv.reader(open(f'/tmp/{tag}_raw.csv')))
hdr,units,r=rows[0],rows[1],rows[2]
keep=('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct','sm__pipe_fma_cycles_active.avg.pct','sm__pipe_fmaheavy','sm__pipe_alu_cycles','sm__pipe_xu','sm__pipe_tensor_cycles_active.avg.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','launch__registers','launch__shared_mem_per_block_dynamic','launch__grid_size','launch__block_size','sm__warps_active.avg.pct','smsp__average_warps_issue_stalled','lts__t_sector_hit_rate.pct','Kernel Name','sm__inst_executed_pipe_lsu','smsp__inst_executed_pipe')
out=open(f'gpurun_out/{tag}_summary.csv','w')
out.write('metric,value,unit\n')
for i,h in enumerate(hdr):
    if any(h.startswith(k) for k in keep): out.write(f'{h},{r[i]},{units[i]}\n')
rows=list(csv.reader(open(f'/tmp/{tag}_sass.csv')))
hdr=rows[1]; body=rows[2:]
ci=hdr.index('Instructions Executed'); si=hdr.index('# Samples')
def val(r,i):
    try: return float(r[i].replace(',',''))
    except: return 0.0
tot=sum(val(r,ci) for r in body); tots=sum(val(r,si) for r in body)
hi=collections.Counter(); hs=collections.Counter()
for r in body:
    t=r[1].strip().split()
    if not t: continue
    op=t[1] if t[0].startswith('@') and len(t)>1 else t[0]
    op='.'.join(op.split('.')[:2])
    hi[op]+=val(r,ci); hs[op]+=val(r,si)
out.write('\nopcode,pct_inst,pct_samples\n')
for op,v in hi.most_common(30): out.write(f'{op},{100*v/tot:.2f},{100*hs[op]/max(tots,1):.2f}\n')
out.write(f'\ntotal_warp_inst,{tot},samples,{tots}\n\nidx,sass,inst_executed,pct_inst,samples,pct_samples\n')
w=csv.writer(out)
for i,r in enumerate(body):
    v=val(r,ci); sm=val(r,si)
    if sm/max(tots,1) > 0.006: w.writerow([i, r[1].strip(), int(v), round(100*v/tot,2), int(sm), round(100*sm/max(tots,1),2)])
out.close()
PY
done
ls -la gpurun_out/*_summary.csv
