CMD="python bench.py --steps 3 --warmup 3 --settle 30 --no-cpu-baseline --no-e2e"
for m in sm grid; do
PBF_TILES=$m ncu --set full --clock-control none -k regex:"lambda_list|delta_list" -s 200 -c 2 -o gpurun_out/tiles_$m -f $CMD > gpurun_out/tiles_$m.log 2>&1
done
ls -la gpurun_out/tiles_*
