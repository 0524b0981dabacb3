fmt='
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()); continue
    print(d["grid"],d["dtype"],"nm",d["nm"],d["tk"],d["ti"],d["jlen"],d["pf"],"stress",d["ms_stress"],"vel",d["ms_vel"],"step",d["ms_step"],"Gc/s",d["gcells_s"],"GB/s",d["GBs"])
'
for lib in h0 h1 fmad; do
  echo "== $lib"
  SWPC3D_LIB=openswpc_b200/lib/libswpc3d_b200_$lib.so python scripts/perf_probe.py --nx 512 --ny 512 --nz 256 --steps 5 --configs "32,8,32,1;32,8,16,1;32,8,64,1;32,8,128,1;64,4,32,1" 2>&1 | python -c "$fmt"
done
echo "== h1 f32"
SWPC3D_LIB=openswpc_b200/lib/libswpc3d_b200_h1.so python scripts/perf_probe.py --nx 512 --ny 512 --nz 256 --steps 5 --dtype f32 --configs "32,8,32,1;64,4,32,1" 2>&1 | python -c "$fmt"
echo "== h1 nm0"
SWPC3D_LIB=openswpc_b200/lib/libswpc3d_b200_h1.so python scripts/perf_probe.py --nx 512 --ny 512 --nz 512 --nm 0 --steps 5 --configs "32,8,32,1;64,4,32,1" 2>&1 | python -c "$fmt"
echo "== h1 big"
SWPC3D_LIB=openswpc_b200/lib/libswpc3d_b200_h1.so python scripts/perf_probe.py --nx 1024 --ny 1024 --nz 512 --steps 3 --configs "32,8,32,1;64,4,32,1" 2>&1 | python -c "$fmt"
