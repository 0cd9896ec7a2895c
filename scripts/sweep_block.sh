for v in 256_3 256_4 128_6 128_7 128_8 256_2 512_1; do
  USV_B200_LIB=$PWD/omniisaacgymenvs_loop_b200/lib/libusv_v_$v.so python bench.py --no-cpu-baseline --no-extra --steps 1000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step']*1e3,2), 'us', round(d['roofline']['frac'],4))"
done
