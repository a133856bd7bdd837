"""Summarise an `ncu --set full` capture of one UNet-WS micro-batch (12 launches: e11 + 11 tensor-core layers) into
profiles/<tag>_ncu_chain.md (table) and profiles/<tag>_ncu_chain.json (what bench.py cites as `roofline.traffic` /
`tensor_pipe_active_pct_ncu`). Runs here (no GPU): it only reads the report.
Usage: python tools/ncu_chain_summary.py gpurun_out/<capture>.ncu-rep <tag> <images> "<command line of the capture>" """
import csv
import io
import json
import subprocess
import sys

LAYERS = ['e11', 'e12', 'e21', 'e22', 'e31', 'e32', 'upconv3', 'd31', 'd32', 'upconv4', 'd41', 'd42']
GF = {'e11': 0.302, 'e12': 19.327, 'e21': 9.664, 'e22': 19.327, 'e31': 9.664, 'e32': 19.327, 'upconv3': 4.295,
      'd31': 38.655, 'd32': 19.327, 'upconv4': 4.295, 'd41': 38.655, 'd42': 19.361}
M = {'name': 'Kernel Name', 'ms': 'gpu__time_duration.sum', 'tensor': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
     'l1tex': 'l1tex__throughput.avg.pct_of_peak_sustained_active', 'rd': 'dram__bytes_read.sum', 'wr': 'dram__bytes_write.sum',
     'lts': 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'regs': 'launch__registers_per_thread',
     'ghz': 'sm__cycles_elapsed.avg.per_second', 'smem': 'launch__shared_mem_per_block_dynamic'}


def main():
    rep, tag, images, cmd = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    terms = json.loads(sys.argv[5]) if len(sys.argv) > 5 else {}
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {k: hdr.index(v) for k, v in M.items() if v in hdr}

    def scale(k, v):   # to GB / ms regardless of the unit ncu picked
        u = units[col[k]].lower()
        f = float(v.replace(',', ''))
        if k in ('rd', 'wr'):
            return f * {'gbyte': 1.0, 'mbyte': 1e-3, 'kbyte': 1e-6, 'byte': 1e-9}[u]
        if k == 'ms':
            return f * {'ms': 1.0, 'us': 1e-3, 'msecond': 1.0, 'usecond': 1e-3, 'second': 1e3, 'ns': 1e-6, 'nsecond': 1e-6}[u]
        return f

    assert len(data) >= len(LAYERS), f'{len(data)} launches captured, expected {len(LAYERS)}'
    out = []
    for layer, r in zip(LAYERS, data[:len(LAYERS)]):
        d = {k: (r[c] if k == 'name' else scale(k, r[c])) for k, c in col.items()}
        kern = d['name'].split('::')[-1].split('(')[0]
        t = terms.get(layer, 3) if layer != 'e11' else 1
        tf = GF[layer] * images / d["ms"]   # GFLOP per ms = TFLOP/s
        out.append({'layer': layer, 'kernel': kern, 'ms': round(d['ms'], 4), 'tensor_active_pct': round(d['tensor'], 1),
                    'algorithmic_tflops': round(tf, 1), 'issued_tflops': round(tf * t, 1), 'terms': t, 'dram_read_gb': round(d['rd'], 3),
                    'dram_write_gb': round(d['wr'], 3), 'l1tex_pct': round(d['l1tex'], 1), 'lts_pct': round(d['lts'], 1),
                    'regs': int(d['regs']), 'sm_ghz': round(d['ghz'], 3)})
    total_ms = sum(o['ms'] for o in out)
    traffic = sum(o['dram_read_gb'] + o['dram_write_gb'] for o in out)
    js = {'captured': tag, 'command': cmd, 'images': images, 'layers': out, 'total_ms': round(total_ms, 3),
          'dram_bytes_per_image': traffic * 1e9 / images,
          'tensor_active_pct': {o['layer']: o['tensor_active_pct'] for o in out if o['layer'] != 'e11'},
          'note': 'ncu serialises launches and replays them (cold caches): shares of the step and per-kernel counters are evidence, absolute times are not bench values'}
    json.dump(js, open(f'profiles/{tag}_ncu_chain.json', 'w'), indent=1)
    with open(f'profiles/{tag}_ncu_chain.md', 'w') as f:
        f.write(f'# ncu --set full of one UNet-WS micro-batch ({images} images 512x512), {tag}\n\nCommand (after the same command exited 0 without ncu):\n`{cmd}`\n\n')
        f.write('| layer | kernel | ms | share | tensor pipe active % | MMAs per MAC | algorithmic TFLOP/s | issued TFLOP/s | DRAM read GB | DRAM write GB | l1tex % | L2 % | regs | SM GHz |\n|' + '---|' * 14 + '\n')
        for o in out:
            f.write(f"| {o['layer']} | {o['kernel']} | {o['ms']:.3f} | {100 * o['ms'] / total_ms:.1f} % | {o['tensor_active_pct']} | {o['terms']} | {o['algorithmic_tflops']:.0f} | "
                    f"{o['issued_tflops']:.0f} | {o['dram_read_gb']:.2f} | {o['dram_write_gb']:.2f} | {o['l1tex_pct']} | {o['lts_pct']} | {o['regs']} | {o['sm_ghz']} |\n")
        f.write(f'\nSum: {total_ms:.3f} ms per {images} images under ncu (cold, serialised) = {images / total_ms * 1e3:.0f} images/s; '
                f'DRAM traffic {traffic:.1f} GB = {traffic / images:.3f} GB per image.\n')
    print(open(f'profiles/{tag}_ncu_chain.md').read())


if __name__ == '__main__':
    main()
