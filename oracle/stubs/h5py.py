"""TEST INFRASTRUCTURE ONLY -- stand-in so that the reference package imports without h5py (not on the decoding path)."""
