cd /root/repo
M=gpu__time_duration.sum,smsp__inst_executed.sum
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/launches_resident.csv -k regex:'k_gather|k_units|k_deblock|k_sao|k_pad|k_ingest|k_pack|k_coeff' python tools/resident_probe.py > gpurun_out/resident.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/launches_resident.csv') if l.startswith('"'))]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
agg=collections.OrderedDict()
for r in rows[1:]:
    if not r[ix['ID']].isdigit(): continue
    e=agg.setdefault(int(r[ix['ID']]),{"k":r[ix['Kernel Name']].split('(')[0][-30:], "g":r[ix['Grid Size']]})
    e[r[ix['Metric Name']]]=float(r[ix['Metric Value']].replace(',',''))
per=collections.OrderedDict()
for e in list(agg.values())[-60:]:
    p=per.setdefault((e["k"],e["g"]),[0,0,0]); p[0]+=1; p[1]+=e['gpu__time_duration.sum']/1e3; p[2]+=e['smsp__inst_executed.sum']
for k,v in per.items(): print(k, "n=%d avg %.1f us, %.2f M inst" % (v[0], v[1]/v[0], v[2]/v[0]/1e6))
PY
