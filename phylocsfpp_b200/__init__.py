"""phylocsfpp_b200 — B200-native (sm_100a) implementation of PhyloCSF++'s per-codon-column
phylogenetic likelihood hot path.  See DESIGN.md."""
__version__ = "0.1.0"
