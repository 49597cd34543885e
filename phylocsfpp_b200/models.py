"""Model loading: built-in parameter sets, external P.nh / P_coding.ECM / P_noncoding.ECM, --species.

Mirror of the reference's `load_model` (src/models.hpp:1757-1856) and its tables:
  * the 11 built-in models (models.hpp:13-1455) ship as files under data/models/ in the reference's own
    external model format (written by tools/extract_builtin_models.py);
  * `sequence_name_mapping` (models.hpp:1468-1706) ships as data/species_aliases.tsv and can be extended
    with `update_sequence_name_mapping` (models.hpp:1709-1740, the `--mapping` option).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

from . import newick
from .ecm import load_ecm

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def builtin_models() -> List[str]:
    with open(os.path.join(_DATA, "models", "INDEX")) as fh:
        return [ln.strip() for ln in fh if ln.strip()]


def _load_aliases() -> Dict[str, List[str]]:
    out: Dict[str, List[str]] = {}
    with open(os.path.join(_DATA, "species_aliases.tsv")) as fh:
        for ln in fh:
            ln = ln.rstrip("\n")
            if not ln:
                continue
            common, _, alts = ln.partition("\t")
            out[common] = [a for a in alts.split(",") if a]
    return out


sequence_name_mapping: Dict[str, List[str]] = _load_aliases()


def update_sequence_name_mapping(path: str) -> None:
    """models.hpp:1709-1740: two-column TSV `common_name<TAB>assembly_name`."""
    with open(path) as fh:
        for ln in fh:
            parts = ln.split()
            if len(parts) < 2:
                continue
            common, sci = parts[0], parts[1]
            names = sequence_name_mapping.setdefault(common, [])
            if sci not in names:
                names.append(sci)


@dataclass
class Model:
    """struct Model (models.hpp:1742-1755) without the HMM."""
    name: str
    S_c: np.ndarray
    f_c: np.ndarray
    S_nc: np.ndarray
    f_nc: np.ndarray
    root: newick.Node            # phylo_tree (double branch lengths; BLS uses these)
    tree: newick.FlatTree        # phylo_array
    seqid_to_phyloid: Dict[str, int]

    @property
    def nl(self) -> int:
        return self.tree.nl


def load_model(model_name_or_path: str, selected_species: str = "") -> Model:
    if model_name_or_path in builtin_models():
        prefix = os.path.join(_DATA, "models", model_name_or_path)
    else:
        prefix = model_name_or_path
    for suffix in ("_coding.ECM", "_noncoding.ECM", ".nh"):
        if not os.path.exists(prefix + suffix):
            raise FileNotFoundError(
                f"Could not open model file '{prefix + suffix}'. Please pass the prefix to the model files "
                f"without any file endings, or one of: {', '.join(builtin_models())}")
    S_c, f_c = load_ecm(prefix + "_coding.ECM")
    S_nc, f_nc = load_ecm(prefix + "_noncoding.ECM")
    with open(prefix + ".nh") as fh:
        root = newick.parse(fh.read())

    if selected_species:
        labels = {lf.label for lf in newick.leaves(root)}
        selected = set()
        for s in selected_species.split(","):
            s = s.lower()
            if s in labels:
                selected.add(s)
                continue
            found = False
            for common, alts in sequence_name_mapping.items():
                if s in alts:
                    found = True
                    selected.add(common)
            if not found:
                selected.add(s)
        missing = sorted(selected - labels)
        if missing:
            raise ValueError("The following selected species are missing in the phylogenetic tree: "
                             + ", ".join(missing))
        newick.reduce(root, selected)
        assert root.branch_length == 0.0

    tree = newick.flatten(root)
    seqid: Dict[str, int] = {}
    for i, label in enumerate(tree.labels):
        if label:
            seqid.setdefault(label, i)
            for alt in sequence_name_mapping.get(label, []):
                seqid.setdefault(alt.lower(), i)
    return Model(os.path.basename(prefix), S_c, f_c, S_nc, f_nc, root, tree, seqid)
