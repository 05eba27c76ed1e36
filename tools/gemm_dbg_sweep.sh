for d in 0 2 1 5; do for r in 23 0; do echo "DBG=$d RESIDENT=$r"; MAE_GEMM_DBG=$d MAE_GEMM_RESIDENT=$r python tools/gemm_micro.py 4 2>&1 | grep "us"; done; done
