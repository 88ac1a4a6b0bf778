python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -2
for st in 100 100; do
python bench.py --steps 30 --warmup 5 --settle $st --no-cpu-baseline --no-e2e > gpurun_out/tb_$st.json 2>gpurun_out/tb.err
python -c "
import json
d=json.load(open('gpurun_out/tb_$st.json')); print('settle$st', d['ms_per_step'], d['roofline']['ms_per_step_by_family']['lambda'], d['roofline']['ms_per_step_by_family']['delta'])"
done
