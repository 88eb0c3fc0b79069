"""Top source lines by sampled stall count from an .ncu-rep captured with --import-source on."""
import csv, subprocess, sys, collections, io
def main(path, top=25):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    if not out.strip():
        out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    for i, r in enumerate(rows):
        if any('Sampl' in c for c in r):
            hdr = i; break
    if hdr is None:
        print('no sampling columns; header candidates:', rows[:3]); return
    H = rows[hdr]
    print(H)
    si = next(i for i, c in enumerate(H) if c.startswith('# Samples') or 'Sampling Data (All)' in c or c == 'Samples')
    src_i = next((i for i, c in enumerate(H) if c in ('Source', 'CUDA', 'Source (CUDA)')), 1)
    data = []
    for r in rows[hdr + 1:]:
        if len(r) <= si: continue
        try: v = float(r[si].replace(',', ''))
        except ValueError: continue
        data.append((v, r))
    tot = sum(v for v, _ in data) or 1
    for v, r in sorted(data, key=lambda t: -t[0])[:top]:
        print(f'{100*v/tot:5.1f}%  {" | ".join(c[:90] for c in r[:4])}')
if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
