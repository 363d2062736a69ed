for lib in build/libsn_pf3.so build/libsupplynet_head.so build/libsn_pf6.so build/libsupplynet_head.so build/libsn_pf3.so; do
SUPPLYNET_LIB=$lib python bench.py --steps 50 --warmup 5 --no-cpu --no-extra --no-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print('$lib', d['ms_per_step'], round(d['roofline']['frac'],4))"
done
