"""Stub package: load_data_ppi.py:6 imports matplotlib.pyplot for a plotting helper the training path never calls."""
