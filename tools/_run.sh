for lib in libmcalf_b200.so libexp_nosleep.so libexp_barrier.so; do
echo "== $lib"
MCALF_B200_LIB=$PWD/mc-alf_b200/$lib python tools/exp_rates.py
done
