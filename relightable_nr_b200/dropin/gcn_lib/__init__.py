"""Drop-in for the reference's ``gcn_lib`` package (dense variant: the only one network.py imports, network.py:7)."""
