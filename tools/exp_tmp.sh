run() { echo "== $1"; shift; env "$@" timeout 300 python bench.py --short --steps 40 --warmup 5 2>&1 | grep -o '"ms_per_step": [0-9.]*\|rror.*' | head -3; }
run "128thr default (TPP2, 4 CTA)" A=1
run "256thr TPP4" ERIRT_B200_LIB=$PWD/lib256_tmp.so ERIRT_TPP=4
run "256thr TPP4 2cta" ERIRT_B200_LIB=$PWD/lib256_tmp.so ERIRT_TPP=4 ERIRT_CTAS_PER_SM=2
run "256thr TPP2" ERIRT_B200_LIB=$PWD/lib256_tmp.so ERIRT_TPP=2
