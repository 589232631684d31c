import json, sys
d=json.load(open(sys.argv[1]))
print('value %.0f  t_dp %.3f ms  dtype %s  launches %s' % (d['value'], d['ms_per_step'], d['dtype'], d.get('gpu_launches_per_step')))
print('e2e %.0f  %.3f ms' % (d['e2e']['value'], d['e2e']['ms_per_step']))
print('kernels', d['kernel_ms_per_step'])
v=d.get('verify') or {}
print('verify', v.get('ok'), {k:(round(x,6) if isinstance(x,float) else x) for k,x in v.items() if 'err' in k})
r=d['roofline']; print('roofline', {k:r[k] for k in ('achieved','peak','frac','contract_ms_per_step')})
print('clocks', d.get('clocks'))
for k,v in (d.get('other_configs') or {}).items():
    if isinstance(v,dict):
        print(k, {kk:(round(vv,4) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','samples_per_s','graph_ms_per_step','GB/s','frac_of_measured_hbm_peak','kernel_ms_per_step')}, 'e2e_ms', (v.get('e2e') or {}).get('ms_per_step'))
    else: print(k, v)
print('bw', {k:round(v['frac_of_measured_hbm_peak'],3) for k,v in (d.get('bandwidth_kernels') or {}).items()})
print('cpu', (d.get('cpu_baseline') or {}).get('value'))
