#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step, DRAM bytes of every GEMM launch, full captures of the top kernels
tag=${1:-r02}
out=gpurun_out; mkdir -p $out
# every launch of one step (the step is 76 kernel launches + the scoring tail; 2 warm-up steps skipped by -s)
ncu --metrics gpu__time_duration.sum --clock-control none -s 152 -c 76 --csv --log-file $out/${tag}_launches.csv python tools/ncu_target.py 3 > /dev/null 2>&1
python tools/agg_launches.py $out/${tag}_launches.csv > $out/${tag}_launches_summary.txt 2>&1; head -14 $out/${tag}_launches_summary.txt
# DRAM bytes of every GEMM launch of one step
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tc -s 55 -c 55 --csv --log-file $out/${tag}_gemm_dram.csv python tools/ncu_target.py 2 > /dev/null 2>&1
python - <<P
import csv
rows=[l for l in open("$out/${tag}_gemm_dram.csv") if not l.startswith("==")]
tot={}
for r in csv.DictReader(rows):
    v=float(r["Metric Value"].replace(",","")); u=r["Metric Unit"]; n=r["Metric Name"]
    if "bytes" in n:
        v*= {"byte":1,"Kbyte":1e3,"Mbyte":1e6,"Gbyte":1e9}.get(u,1)
        tot[n]=tot.get(n,0)+v
print({k:round(v/1e9,3) for k,v in tot.items()}, "GB over the GEMM launches of one step")
P
# full captures: second step's conv1, qkv0 / out0 / fc1_0 / fc2_0, and attention / conv0 / posconv
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 55 -c 1 -o $out/${tag}_ncu_gemm_conv1 python tools/ncu_target.py 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 62 -c 4 -o $out/${tag}_ncu_gemm_layer0 python tools/ncu_target.py 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"attention_fa|conv0_mma|posconv_kernel" -s 3 -c 3 -o $out/${tag}_ncu_misc python tools/ncu_target.py 2 > /dev/null 2>&1
ls -la $out | grep ${tag}_ | awk '{print $5, $9}'
