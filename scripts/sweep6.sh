fmt='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()); continue
    print(d["grid"],d["dtype"],"nm",d["nm"],d["tk"],d["ti"],d["jlen"],d["pf"],"stress",d["ms_stress"],"vel",d["ms_vel"],"step",d["ms_step"],"Gc/s",d["gcells_s"],"GB/s",d["GBs"])
'
python scripts/dbg_tma.py a b c d 2>&1 | tail -4
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python scripts/perf_probe.py --nx 1024 --ny 1024 --nz 512 --steps 3 --configs "32,8,16,1,32,2;32,8,16,1,32,1;32,8,16,1,16,2;32,8,16,1,64,2" 2>&1 | python -c "$fmt"
