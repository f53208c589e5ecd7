for g in "512 512 512" "256 256 256" "128 128 256"; do python tools/poisson_only.py $g 5; done
python -m pytest tests -m gpu -q -x -k "poisson" 2>&1 | tail -3
