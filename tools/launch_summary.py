"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import csv, sys, collections
def main(path):
    rows = [r for r in csv.reader(open(path, errors='ignore')) if r and not r[0].startswith('==')]
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    H = rows[hdr]
    kn, mv, mu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    d = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= mv: continue
        v = float(r[mv].replace(',', ''))
        if r[mu] in ('ns', 'nsecond'): v /= 1e3
        elif r[mu] in ('ms', 'msecond'): v *= 1e3
        name = r[kn].split('(')[0].replace('tscm::', '')
        d.setdefault(name, []).append(v)
    iter_k = [k for k in d if not any(s in k for s in ('k_init', 'k_prep', 'k_jacobi', 'k_transpose', 'k_dfma', 'k_eval_rows'))]
    tot = sum(sum(d[k]) / len(d[k]) for k in iter_k)
    print(f'{"kernel":42s} {"launches":>8s} {"mean us":>9s}  share of iteration')
    for k in sorted(d, key=lambda k: -sum(d[k]) / len(d[k])):
        m = sum(d[k]) / len(d[k])
        share = f'{100 * m / tot:5.1f}%' if k in iter_k else '(setup / inspection)'
        print(f'{k:42s} {len(d[k]):8d} {m:9.1f}  {share}')
    print(f'\nsum of the kernels of one LM iteration: {tot:.1f} us')
if __name__ == '__main__':
    main(sys.argv[1])
