for g in "512 512 512" "256 256 256"; do
echo "== default $g"; python tools/poisson_only.py $g 5
echo "== MINB=2 stage"; SOPHT_P2_MINB=2 SOPHT_P2_STAGE=1 python tools/poisson_only.py $g 5
echo "== MINB=2 nostage"; SOPHT_P2_MINB=2 SOPHT_P2_STAGE=0 python tools/poisson_only.py $g 5
echo "== MINB=1 nostage"; SOPHT_P2_MINB=1 SOPHT_P2_STAGE=0 python tools/poisson_only.py $g 5
echo "== MINB=1 stage"; SOPHT_P2_MINB=1 SOPHT_P2_STAGE=1 python tools/poisson_only.py $g 5
done
