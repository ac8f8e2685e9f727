"""Summarise an ncu launch list (csv from `ncu --metrics gpu__time_duration.sum --csv`) per kernel."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get('Metric Name') == 'gpu__time_duration.sum':
            v = float(row['Metric Value'].replace(',', ''))
            unit = row['Metric Unit']
            v = v / 1000 if unit == 'ns' else v * 1000 if unit == 'ms' else v
            agg[row['Kernel Name'].split('(')[0][-48:]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':50s} {'launches':>8s} {'mean us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:50s} {len(v):8d} {sum(v) / len(v):10.1f} {sum(v) / tot * 100:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
