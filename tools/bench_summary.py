import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    bk=d['config'].get('by_kind') or d.get('roofline',{}).get('by_kind') or {}
    print(f, 'value %.1f e2e %.1f b1 %.1f frac %.4f' % (d['value'], d['e2e']['value'], d.get('batch1',{}).get('value',0), d['roofline']['frac']), 'vox+spconv ms/scene', d.get('voxelize_spconv_ms_per_scene'))
    def find(o):
        if isinstance(o,dict):
            if 'by_kind' in o: return o['by_kind']
            for v in o.values():
                r=find(v)
                if r: return r
        return None
    bk=find(d)
    if bk: print('   ', {k:(round(v['ms'],2), round(v['tflops'],1)) for k,v in bk.items()})
