import json, glob, sys, os
d = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/r2e'
for f in sorted(glob.glob(d + '/*.log')):
    for ln in open(f):
        if not ln.startswith('{'):
            continue
        j = json.loads(ln)
        if 'value' in j:
            e = j.get('e2e', {})
            rl = j.get('rel_l2') or {}
            print('%-22s N=%d %s value %.4g pts/s  %.3f ms/step | e2e %.3f ms (pinned %.3f) | rel_l2 %s %s | %s' % (
                os.path.basename(f), j['n_gpus'], j['scaling'], j['value'], j['ms_per_step'], e.get('ms_per_step', 0), e.get('caller_pinned_ms_per_step', 0),
                rl.get('trafo'), rl.get('adjoint'), j['implementation']['multi_gpu'][-70:]))
        elif 'ms_pair' in j:
            print('%-22s P=%d group e2e: trafo %.2f adjoint %.2f pair %.2f ms  %.4g pts/s first %.2fs rel %s' % (
                os.path.basename(f), j['n_gpus'], j['ms_trafo'], j['ms_adjoint'], j['ms_pair'], j['points_per_s'], j['first_call_s'], j.get('rel_l2')))
        else:
            print(os.path.basename(f), {k: v for k, v in j.items() if k not in ('workers', 'config')})
