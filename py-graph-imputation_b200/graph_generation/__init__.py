"""Offline data-format helpers that sit in front of the table build (the reference keeps them in
graph_generation/): here only produce_hpf, the README's first step."""
