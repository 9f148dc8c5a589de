"""Mirrors of py5gphy/nr_pusch/nr_ulsch.py and nr_ulsch_decode.py on the batched CUDA chain."""
