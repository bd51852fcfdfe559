run() {
timeout 300 python bench.py --no-cpu --no-e2e $2 > gpurun_out/b_x.json 2>gpurun_out/b_x.err || tail -5 gpurun_out/b_x.err; python -c "
import json; d=json.load(open('gpurun_out/b_x.json')); print('$1', round(d['value']), round(d['ms_per_step'],3), d['config'].get('launch'), {k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"
}
timeout 300 python -m pytest tests -m gpu -x -q -k "graphed or register_from_host" 2>&1 | tail -3
run graph1; run graph2; run eager --no-graph; run graph3
