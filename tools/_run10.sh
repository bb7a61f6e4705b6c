mkdir -p gpurun_out/r2m
for occ in ${OCCS:-1}; do
NG_OCC=$occ timeout 600 ncu --set full --clock-control none --import-source on -k regex:ng_pipe -c 1 -s 1 -o gpurun_out/r2m/ng${TAG}_occ$occ -f python tools/ng_prof.py > gpurun_out/r2m/ncu_ng$occ.log 2>&1; echo "ncu occ $occ rc=$?"
done
