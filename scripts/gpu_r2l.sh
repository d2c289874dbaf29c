#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; rm -f $OUT/r2l_misc_launches.txt
for c in "44100 47999.5 VHQ float32 256" "47999 44100 HQ float32 256" "48000 44100 LQ float32 256" "48000 44100 MQ float32 256" "48000 44100 QQ float32 256" "44100 48000 VHQ float64 256" "48000 8000 HQ float32 256"; do
  echo "== $c" >> $OUT/r2l_misc_launches.txt
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/l.csv python scratch/gpu_misc_one.py $c 2>&1 | tail -1 >> $OUT/r2l_misc_launches.txt
  python - >> $OUT/r2l_misc_launches.txt <<'PY'
import csv
rows=list(csv.reader(open('/tmp/l.csv')))
hdr=[r for r in rows if r and r[0]=='ID'][0]
ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); ig=hdr.index('Grid Size'); ib=hdr.index('Block Size')
data=[r for r in rows if len(r)==len(hdr) and r[0]!='ID']
n=len(data)//3
for r in data[-n:]: print('   %-60s grid %-18s block %-14s %8.1f us' % (r[ik][:60], r[ig], r[ib], float(r[iv])/1e3))
PY
done
cat $OUT/r2l_misc_launches.txt
