mkdir -p gpurun_out
CMD="python tools/one_search.py 65536 96 8192"
$CMD > gpurun_out/plain.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"nearest" -s 5 -c 6 --csv --log-file gpurun_out/search_launches.csv $CMD > gpurun_out/ncu_search.log 2>&1
python - <<EOF
import csv
rows=list(csv.reader(open("gpurun_out/search_launches.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
H=rows[h]
for r in rows[h+1:]:
    print(r[H.index("Kernel Name")][:30], r[H.index("Metric Value")])
EOF
