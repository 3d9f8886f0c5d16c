"""alignTools DP core, B200-native.  The product lives in aligntools.c_b200."""
