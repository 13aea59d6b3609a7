"""PDB text writer with the reference's entry point (pepflow/modules/protein/writers.py:10-83, called from
models_con/sample.py:65-66,89-94,115-120) - the on-disk format after the sampling path.

The reference builds a Biopython structure and lets Bio.PDB.PDBIO serialise it; Biopython is not in this image, so
the records are formatted here directly in PDBIO's published fixed-column layout (ATOM / TER / END, serial numbers
restarting at 1, one TER per chain, element = first letter of the atom name, occupancy 1.00, B-factor 0.00).
Parity with PDBIO's bytes is UNPINNED (nothing to run it against offline); tests check the column layout and a
write -> parse round trip.  Host-side Python by nature: it formats a few thousand lines once per sample."""
import torch

from .constants import AA, heavyatom_names

_ATOM = "ATOM  %5i %-4s%c%3s %c%4i%c   %8.3f%8.3f%8.3f%6.2f%6.2f      %4s%2s%2s\n"
_TER = "TER   %5i      %3s %c%4i%c" + " " * 53 + "\n"   # 80 columns


def _fullname(name):
    # writers.py:54-57: element symbol in columns 13-14 for one-letter elements
    return {1: " %s  ", 2: " %s ", 3: " %s"}.get(len(name), "%s") % name


def _select(v, mask):
    if isinstance(v, str):
        return "".join(s for i, s in enumerate(v) if mask[i])
    if isinstance(v, (list, tuple)):
        return [s for i, s in enumerate(v) if mask[i]]
    if isinstance(v, torch.Tensor):
        return v[mask]
    return v


def pdb_string(data):
    """data: chain_nb [L] i64, aa [L] i64, pos_heavyatom [L,A,3], mask_heavyatom [L,A] bool, chain_id / icode (str or
    list, one entry per residue), resseq [L] i64 - the dict the reference's save_pdb takes."""
    chain_nb = torch.as_tensor(data["chain_nb"]).cpu()
    aa_all = torch.as_tensor(data["aa"]).cpu()
    pos_all = torch.as_tensor(data["pos_heavyatom"]).cpu().double()
    mask_all = torch.as_tensor(data["mask_heavyatom"]).cpu().bool()
    resseq_all = torch.as_tensor(data["resseq"]).cpu()
    lines, serial = [], 1
    for ch in chain_nb.unique().tolist():
        sel = chain_nb == ch
        aa, pos, mask, resseq = aa_all[sel], pos_all[sel], mask_all[sel], resseq_all[sel]
        chain_id, icode = _select(data["chain_id"], sel), _select(data["icode"], sel)
        cid, last = chain_id[0], None
        for r in range(aa.shape[0]):
            resname = AA(int(aa[r])).name
            names = heavyatom_names(AA(int(aa[r])))
            for i, name in enumerate(names[:pos.shape[1]]):
                if name == "" or not bool(mask[r, i]):
                    continue
                x, y, z = pos[r, i].tolist()
                lines.append(_ATOM % (serial, _fullname(name), " ", resname, cid, int(resseq[r]), icode[r], x, y, z,
                                      1.0, 0.0, "    ", name[0].rjust(2), "  "))
                serial += 1
                last = (resname, int(resseq[r]), icode[r])
        if last is not None:
            lines.append(_TER % (serial, last[0], cid, last[1], last[2]))
            serial += 1
    lines.append("END   \n")
    return "".join(lines)


def save_pdb(data, path=None):
    """Writes `data` as a PDB file (when `path` is given) and returns the text."""
    text = pdb_string(data)
    if path is not None:
        with open(path, "w") as f:
            f.write(text)
    return text


def parse_pdb_atoms(text):
    """(serial, name, resname, chain, resseq, xyz) per ATOM record - the inverse used by the round-trip test."""
    out = []
    for ln in text.splitlines():
        if ln.startswith("ATOM  "):
            out.append((int(ln[6:11]), ln[12:16].strip(), ln[17:20], ln[21], int(ln[22:26]),
                        (float(ln[30:38]), float(ln[38:46]), float(ln[46:54]))))
    return out
