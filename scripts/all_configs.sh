cd /root/repo
for c in config0 config1 config2 config2x4 config2x16 config3; do
  python bench.py --config $c --no-cpu-baseline --steps 2 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$c', '%.1f M samples/s'%(d['value']/1e6), 'e2e %.1f'%(d['e2e']['value']/1e6), 'ms/step %.1f'%d['ms_per_step'], d['kernel_ms_first_batches'], 'rays/sample %.2f'%d['rays_per_sample'])"
done
python bench.py --config config4 --no-cpu-baseline --steps 1 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('config4', '%.1f M samples/s'%(d['value']/1e6), 'e2e %.1f'%(d['e2e']['value']/1e6), 'ms/step %.1f'%d['ms_per_step'], d['kernel_ms_first_batches'])"
