import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('value %.4g e2e %.4g frac %.4f clocks %s'%(d['value'], d['e2e']['value'], d['roofline']['whole_pass']['frac_of_fma_peak'], d['clocks']))
for k,v in d['roofline']['per_kernel'].items(): print('  ',k, round(v['ms_per_launch'],4), round(v['frac_of_fma_peak'],4))
