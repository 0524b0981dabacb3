for lib in minb2 minb3 minb4; do
  echo "== $lib"
  SWPC3D_LIB=openswpc_b200/lib/libswpc3d_b200_$lib.so python scripts/perf_probe.py --nx 512 --ny 512 --nz 256 --steps 5 --configs "32,8,32,0;32,8,32,1;32,8,32,2;64,4,32,1;64,4,64,2;128,2,32,1" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()); continue
    print(d['tk'],d['ti'],d['jlen'],d['pf'],'stress',d['ms_stress'],'vel',d['ms_vel'],'step',d['ms_step'],'Gc/s',d['gcells_s'],'GB/s',d['GBs'])
"
done
