cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -x -q 2>&1 | tail -3
for wl in short8 large sweep long16; do
  echo "$wl w=100: $(python scripts/packed_probe.py $wl --only-resident --steps 5 2>&1 | tail -1)"
done
echo "sweep w=500: $(python scripts/packed_probe.py sweep --only-resident --steps 5 --w 500 2>&1 | tail -1)"
echo "sweep w=32: $(python scripts/packed_probe.py sweep --only-resident --steps 5 --w 32 2>&1 | tail -1)"
echo "small: $(python scripts/packed_probe.py small --only-resident --steps 5 2>&1 | tail -1)"
