#!/bin/bash
# Turn the gpurun_out/ artefacts of `tools/gpu_round.sh <tag>` into the committed summaries under profiles/ (run where ncu is).
# usage: tools/make_profiles.sh <tag> <round>   e.g. tools/make_profiles.sh r02k r02
tag=$1; rnd=${2:-r02}
B=262144
kms=$(python -c "import json; print(json.load(open('gpurun_out/bench_$tag.json'))['roofline']['kernel_ms'])")
python tools/ncu_summary.py gpurun_out/prof_$tag.ncu-rep profiles/ncu_fast_kernel_$rnd.txt $B > /dev/null
python tools/ncu_opcodes.py gpurun_out/prof_$tag.ncu-rep profiles/ncu_opcodes_$rnd.csv profiles/ncu_opcodes_$rnd.json $B $kms > /dev/null
python - <<PY
import json
p='profiles/ncu_opcodes_$rnd.json'; d=json.load(open(p)); d['source']='profiles/ncu_opcodes_$rnd.csv (ncu --set full --clock-control none --import-source on, python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-strong; third launch)'; d['kernel_ms_source']='CUDA events of the bench run without ncu (gpurun_out/bench_$tag.json)'; json.dump(d,open(p,'w'),indent=1)
PY
for cfg in 2 3; do
  n=$(python -c "print({2:131072,3:65536}[$cfg])")
  python tools/ncu_summary.py gpurun_out/prof_cfg${cfg}_$tag.ncu-rep profiles/ncu_fast_kernel_${rnd}_cfg$cfg.txt $n > /dev/null
  python tools/ncu_opcodes.py gpurun_out/prof_cfg${cfg}_$tag.ncu-rep /tmp/op_cfg$cfg.csv profiles/ncu_opcodes_${rnd}_cfg$cfg.json $n > /dev/null
done
cp gpurun_out/launches_$tag.csv profiles/launches_$rnd.csv
cp gpurun_out/bench_$tag.json profiles/bench_${rnd}_1gpu.json
python - <<PY
import csv, io, json, subprocess
rep='gpurun_out/prof_$tag.ncu-rep'
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out))); h,u,v=rows[0],rows[1],rows[2]
g=lambda k: float(v[h.index(k)].replace(',',''))
scale={'Mbyte':1e6,'Kbyte':1e3,'Gbyte':1e9,'byte':1.0}
b=lambda k: g(k)*scale[u[h.index(k)]]
oc=json.load(open('profiles/ncu_opcodes_$rnd.json'))
d={'kernel':'mcalf_fast_kernel','batch':$B,'dram_bytes_read':b('dram__bytes_read.sum'),'dram_bytes_write':b('dram__bytes_write.sum'),
   'source':'profiles/ncu_fast_kernel_$rnd.txt (ncu --set full --clock-control none, python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep --no-strong)',
   'issue_active_pct':g('smsp__issue_active.avg.pct_of_peak_sustained_active'),
   'pipe_fma_cycles_active_pct':g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
   'pipe_fma_pct':g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
   'pipe_alu_pct':g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
   'pipe_xu_pct':g('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
   'pipe_lsu_pct':g('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'),
   'warp_instructions_per_logL':oc['warp_instructions_per_logL']}
json.dump(d,open('profiles/ncu_traffic.json','w'),indent=1)
print(d)
PY
# SASS of the shipped hot kernel
cuobjdump -sass mc-alf_b200/libmcalf_b200.so > /tmp/all.sass
# the instantiation the bench workload runs: <STATS=false, EXTRAS=false, DENSE=true, ONE_EACH=false> (48 registers, five CTAs per SM at cfg 4)
L=$(grep -n "Function :" /tmp/all.sass | grep "ILb0ELb0ELb1ELb0E" | cut -d: -f1); N=$(grep -n "Function :" /tmp/all.sass | awk -F: -v l=$L '$1>l{print $1; exit}')
[ -z "$N" ] && N=$(wc -l < /tmp/all.sass)
( echo "SASS of mcalf_fast_kernel<STATS=false, EXTRAS=false, DENSE=true, ONE_EACH=false> (sm_100a) from mc-alf_b200/libmcalf_b200.so, cuobjdump -sass; opcode totals first"; 
  sed -n "${L},${N}p" /tmp/all.sass | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s*\/\*.*$//' > /tmp/hot.sass
  awk '{op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); c[op]++} END{for(k in c) printf "%6d %s\n", c[k], k}' /tmp/hot.sass | sort -rn | head -40
  echo; cat /tmp/hot.sass ) > profiles/sass_fast_kernel_$rnd.txt
wc -l profiles/sass_fast_kernel_$rnd.txt
