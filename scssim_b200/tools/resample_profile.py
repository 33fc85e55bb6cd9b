"""Derive a profile with a different read length from a shipped one by nearest-bin resampling.

Read length is a property of the .profile, not a CLI flag (/root/reference/lib/profile/Profile.cpp:976-999);
the shipped HiSeq2500 profile is 125 bp. BASELINE.json's headline config asks for "PE150 with the HiSeq2500
profile", so the bench derives a 150-bin profile from it: row j of the new table = row floor(j*old/new) of
the old one, for the per-bin substitution (read 1 and read 2) and quality tables; everything else is copied.
The result is a valid .profile the reference itself accepts.
"""
import sys


def resample(src_path: str, dst_path: str, new_len: int) -> None:
    lines = open(src_path).read().split("\n")
    out, i, old = [], 0, None

    def take_rows(n_old, n_new):
        nonlocal i
        rows = lines[i:i + n_old]
        i += n_old
        return [rows[j * n_old // n_new] for j in range(n_new)]

    while i < len(lines):
        ln = lines[i]
        if ln.startswith("readLength:"):
            old = int(ln.split(":")[1]); out.append(f"readLength: {new_len}"); i += 1
        elif ln.startswith("binCount:"):
            out.append(f"binCount: {new_len}"); i += 1
        elif ln.startswith("kmer:") and not ln.split(":")[1].strip().isdigit():   # a k-mer block: 2*old rows (read 1 then read 2)
            out.append(ln); i += 1
            out += take_rows(old, new_len); out += take_rows(old, new_len)
        elif ln.startswith("basePairIndx:"):
            out.append(ln); i += 1
            out += take_rows(old, new_len)
        else:
            out.append(ln); i += 1
    open(dst_path, "w").write("\n".join(out))


if __name__ == "__main__":
    resample(sys.argv[1], sys.argv[2], int(sys.argv[3]))
