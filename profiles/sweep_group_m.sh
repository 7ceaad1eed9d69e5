for g in 0 2 4 6 8 12 16; do
echo "GROUP_M=$g"
RDB_OZAKI_GROUP_M=$g timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-extra 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value'],2), round(r['frac'],4), round(r['fp_equivalent']['achieved_tflops'],2), [(s['step'], round(s['ms']/s['count'],3)) for s in r['steps'] if 'gemm' in s['step']])"
done
