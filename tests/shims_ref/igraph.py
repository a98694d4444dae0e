"""Stub: load_data_ppi.py:9 imports igraph for a plotting helper the training path never calls."""
