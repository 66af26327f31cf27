cd $GRAFT_REPO_ROOT
echo "small short: $(python scripts/packed_probe.py small --only-resident --steps 10 2>&1 | tail -1)"
echo "small long:  $(python scripts/packed_probe.py small --only-resident --steps 10 --long 2>&1 | tail -1)"
for n in 2000 5000 20000 40000; do
echo "small n=$n short: $(python scripts/packed_probe.py small --n $n --only-resident --steps 10 2>&1 | tail -1)"
echo "small n=$n long:  $(python scripts/packed_probe.py small --n $n --only-resident --steps 10 --long 2>&1 | tail -1)"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsw_short16 -c 1 -f -o /tmp/prof python scripts/resident_run.py sweep_w100_z100 0 1 > /dev/null 2>&1
python scripts/ncu_summary.py /tmp/prof.ncu-rep 0 --sass > gpurun_out/r02k_sass_sweep_w100_launch0.txt 2>&1
wc -l gpurun_out/r02k_sass_sweep_w100_launch0.txt
