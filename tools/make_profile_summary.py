"""Turn the scratch captures in gpurun_out/ into the committed summaries under profiles/.

    python tools/make_profile_summary.py r01 k_eval4 k_schur_frames k_schur_update k_solve ...

For every kernel: headline ncu metrics + warp-stall breakdown + hottest source lines
(profiles/<round>_<kernel>.txt) and the DRAM traffic of the launch
(profiles/<round>_traffic.json, read by bench.py for roofline.traffic).
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], zip(rows[1], rows[2])))


def to_bytes(val, unit):
    v = float(val.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def main():
    rnd, kernels = sys.argv[1], sys.argv[2:]
    traffic_path = os.path.join(ROOT, 'profiles', f'{rnd}_traffic.json')
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for k in kernels:
        rep = os.path.join(ROOT, 'gpurun_out', f'prof_{k}.ncu-rep')
        if not os.path.exists(rep):
            print('missing', rep); continue
        a = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_summary.py'), rep], capture_output=True, text=True).stdout
        b = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_source.py'), rep, '30'], capture_output=True, text=True).stdout
        b = '\n'.join(l for l in b.splitlines() if not l.startswith("['Line No'"))
        with open(os.path.join(ROOT, 'profiles', f'{rnd}_{k}.txt'), 'w') as f:
            f.write(f'# ncu --set full --clock-control none --import-source on -k regex:{k}  (bench.py --steps 3 --warmup 3)\n')
            f.write('# headline metrics, warp-stall sampling, hottest source lines\n\n')
            f.write(a + '\n--- hottest source lines (share of warp-stall samples) ---\n' + b + '\n')
        m = raw(rep)
        rd = to_bytes(m['dram__bytes_read.sum'][1], m['dram__bytes_read.sum'][0])
        wr = to_bytes(m['dram__bytes_write.sum'][1], m['dram__bytes_write.sum'][0])
        dur = m['gpu__time_duration.sum']
        traffic[k] = {'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes': rd + wr,
                      'duration': f'{dur[1]} {dur[0]}',
                      'fp64_pipe_pct': float(m['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'][1])}
        print(k, traffic[k])
    json.dump(traffic, open(traffic_path, 'w'), indent=1)


if __name__ == '__main__':
    main()
