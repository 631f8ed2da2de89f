"""Per-kernel SASS instruction counts of the in-tree objects -> profiles/r02_sass_evidence.txt (no GPU needed)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = ['# SASS evidence: per-kernel instruction counts from `cuobjdump -sass pylrbms_b200/build/<file>.o` (sm_100a only).',
       '# FP64 has no tcgen05 / UTCxMMA kind: the FP64 tensor pipe on sm_100a is warp-level DMMA.8x8x4 (mma.sync.m8n8k4.f64).',
       '# cp.async.bulk (TMA engine; plain bulk copy, no tensor map) is UBLKCP, its mbarrier SYNCS; cp.async is LDGSTS.',
       '# peer_push_kernel (context.o): multimem.st on the NVSwitch multicast address is a predicated STG.E.128, the stores to the',
       '# mapped NVLink peer pointers are ST.E.128 (generic address space); pcg_iterate_kernel is the cooperative CG kernel.',
       '# Regenerate: python tools/sass_evidence.py', '']
for o in ('band', 'online', 'project', 'context', 'pcg'):
    txt = subprocess.run(['cuobjdump', '-sass', os.path.join(ROOT, 'pylrbms_b200', 'build', o + '.o')], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1); per[fn] = collections.Counter(); continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?', line)
        if m and fn:
            per[fn][m.group(1)] += 1
            if m.group(1) in ('UBLKCP', 'DMMA') and per[fn][m.group(1)] == 1:
                per[fn]['_ex_' + m.group(1)] = re.sub(r'\s+/\* 0x.*', '', line.strip())
    out.append('== %s.o' % o)
    for fn, c in per.items():
        m = re.search(r'\d+([a-z][a-z_0-9]*kernel[a-z_0-9]*)(I[A-Za-z0-9_]*E)?', fn)
        name = (m.group(1) + (' ' + m.group(2) if m.group(2) else '')) if m else fn[-40:]
        keys = [k for k in ('DMMA', 'UBLKCP', 'SYNCS', 'LDGSTS', 'LDS', 'STS', 'STG', 'DFMA', 'SHFL', 'BAR', 'MUFU') if c.get(k)]
        out.append('  %-44s %s' % (name[:44], ' '.join('%s=%d' % (k, c[k]) for k in keys)))
        for k in ('DMMA', 'UBLKCP'):
            if c.get('_ex_' + k):
                out.append('      e.g. ' + c['_ex_' + k])
    out.append('')
open(os.path.join(ROOT, 'profiles', sys.argv[1] if len(sys.argv) > 1 else 'r02_sass_evidence.txt'), 'w').write('\n'.join(out) + '\n')
