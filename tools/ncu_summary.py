"""Read key metrics out of committed .ncu-rep files (ncu -i ... --page raw --csv) into profiles/r0N_ncu_summary.md.
    python tools/ncu_summary.py [r01|r02]"""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.avg.per_second',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum']


ROUNDS = {
    "r01": (("profiles/r01_ffn_gemm.ncu-rep", "grouped FFN GEMMs"),
            ("profiles/r01_router_plan_permute_combine.ncu-rep", "router / plan / permute / combine"),
            ("profiles/r01_router_final.ncu-rep", "router, final version of the round (29 warps, three routing groups)"),
            ("profiles/r01_decode_T8.ncu-rep", "decode-sized call, T = 8 (tools/decode_once.py): fused front end, "
                                               "weight-streaming GEMM-1 / GEMM-2, combine")),
    "r02": (("profiles/r02_ffn_gemm.ncu-rep", "grouped FFN GEMMs"),
            ("profiles/r02_router_plan_permute_combine.ncu-rep", "router / plan / permute / combine"),
            ("profiles/r02_decode_T8.ncu-rep", "decode-sized call, T = 8 (tools/decode_once.py): front end without the row gather, "
                                               "weight-streaming GEMM-1 (token rows by TMA gather4) / GEMM-2, combine")),
}


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = [f"# ncu summary, round {int(rnd[1:])}\n\n",
           "Read with `ncu -i <file>.ncu-rep --page raw --csv` (tools/ncu_summary.py); one launch per kernel, captured after\n"
           "warm-up with `--set full --clock-control none --import-source on` from `bench.py --steps 3 --warmup 3`\n"
           "(BASELINE.json configs[1]: 8 x 2048 tokens, bf16).\n"]
    for f, note in ROUNDS[rnd]:
        txt = subprocess.run(["ncu", "-i", os.path.join(ROOT, f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(txt.splitlines()))
        hdr, units = r[0], r[1]
        out.append(f"\n## {note} (`{f}`)\n")
        for row in r[2:]:
            name = row[hdr.index('Kernel Name')]
            out.append(f"\n### `{name[:80]}`\n\n| metric | value |\n|---|---|\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    out.append(f"| {w} | {row[i]} {units[i]} |\n")
    out.append(open(os.path.join(ROOT, "profiles", "_reading.md")).read())
    open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_summary.md"), "w").write("".join(out))


if __name__ == "__main__":
    main()
