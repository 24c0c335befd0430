"""Test oracle package - see oracle/oracle.py.  Never imported by phylo_utils_b200."""
