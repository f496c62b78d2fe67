for P in 250 150 100 60 40 25 15; do echo "== P=$P"; QK_P=$P QK_T=230 QK_ONLY=1:8,1:4,1:2,1:-1 python scripts/quick_k2.py 2>&1 | grep variant; done
