mkdir -p gpurun_out
for v in "auto:" "subdomain:0,4" "subdomain:1,4" "subdomain:2,4" "subdomain:3,4" "resident:"; do
eng=${v%%:*}; sub=${v##*:}
JJ_ENGINE=$eng JJ_SUBDOMAIN=$sub timeout 300 python tools/config_sweep.py cfg1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', d['engine'], d['device_us_per_time_step'], 'us/step', round(d['junction_steps_per_s_device']/1e9,3), 'G')"
done
