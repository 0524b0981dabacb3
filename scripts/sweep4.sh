fmt='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()); continue
    print(d["grid"],d["dtype"],"nm",d["nm"],d["tk"],d["ti"],d["jlen"],d["pf"],"stress",d["ms_stress"],"vel",d["ms_vel"],"step",d["ms_step"],"Gc/s",d["gcells_s"],"GB/s",d["GBs"])
'
for v in t32x8 t64x4 t128x2; do
echo "== $v"
SWPC3D_LIB=openswpc_b200/lib/libswpc3d_b200_$v.so python scripts/dbg_tma.py a c 2>&1 | tail -2
SWPC3D_LIB=openswpc_b200/lib/libswpc3d_b200_$v.so python scripts/perf_probe.py --nx 1024 --ny 1024 --nz 512 --steps 3 --configs "32,8,16,1,64,1;32,8,16,1,32,1;32,8,16,1,164,1;32,8,16,1,984,1" 2>&1 | python -c "$fmt"
done
