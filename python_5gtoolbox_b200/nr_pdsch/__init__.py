"""Mirrors of py5gphy/nr_pdsch/nr_dlsch.py and nr_dlsch_decode.py on the batched CUDA chain."""
