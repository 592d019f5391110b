for n in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2965$n bench.py --gpus $n --steps 300 --warmup 20 --no-e2e > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; echo "N=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/scale_$n.json')); print('N=$n', d['value'], d['ms_per_step'], d['roofline']['avg_launch_ms'], {k:v.get('ms_per_step') for k,v in d['roofline_extra'].items()}, 'spce', d['spce']['value'], d['spce']['ms_per_step'], {k:v.get('ms_per_step', v.get('rho_plus_force_ms')) for k,v in d['spce']['roofline_extra'].items()})"
done
