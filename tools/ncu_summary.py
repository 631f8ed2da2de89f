"""Turn `ncu --page raw --csv` output into the per-kernel summary lines kept under profiles/ and (with --traffic) into
profiles/traffic.json, the table bench.py reads `roofline.traffic` from.

    python tools/ncu_summary.py raw.csv [--traffic KERNEL_SUBSTRING:NAME:SOURCE_FILE:WORKLOAD ...]
"""
import csv, hashlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__inst_executed_pipe_fp64_op_dmma.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum']


def to_float(x):
    try:
        return float(x.replace(',', ''))
    except Exception:
        return None


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path, newline='')))
    # the header row is the one containing "Kernel Name"; the units row follows
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    head, units, body = rows[h], rows[h + 1], rows[h + 2:]
    ik = head.index('Kernel Name')
    idx = {c: head.index(c) for c in COLS if c in head}
    out = []
    for r in body:
        if len(r) <= ik:
            continue
        rec = {'kernel': r[ik].split('(')[0]}
        for c, i in idx.items():
            rec[c] = (to_float(r[i]), units[i])
        out.append(rec)
    for rec in out:
        t = rec.get('gpu__time_duration.sum', (None, ''))
        print('%-60s' % rec['kernel'][:60], ' '.join('%s=%s%s' % (c.split('.')[0].split('__')[-1][:22], ('%.4g' % v[0]) if v[0] is not None else '-', v[1])
                                                       for c, v in rec.items() if c != 'kernel'))
    traffic = [a for a in sys.argv[2:] if a != '--traffic']
    if traffic:
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        table = json.load(open(tpath)) if os.path.exists(tpath) else {}
        for spec in traffic:
            sub, name, src, workload = spec.split(':', 3)
            sel = [rec for rec in out if sub in rec['kernel']]
            if not sel:
                continue

            def gb(rec, c):
                v, u = rec[c]
                return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(u, 1)
            tot = sum(gb(rec, 'dram__bytes_read.sum') + gb(rec, 'dram__bytes_write.sum') for rec in sel)
            per = tot / len(sel) if name != 'projection_plan' else tot
            with open(os.path.join(ROOT, 'pylrbms_b200', 'csrc', src), 'rb') as f:
                sha = hashlib.sha256(f.read()).hexdigest()
            table[name] = {'dram_bytes_per_launch': per, 'launches_in_capture': len(sel), 'source': src, 'source_sha256': sha,
                           'workload': workload, 'capture': os.path.basename(path)}
        json.dump(table, open(tpath, 'w'), indent=1)


if __name__ == '__main__':
    main()
