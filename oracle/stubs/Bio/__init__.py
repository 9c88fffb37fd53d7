"""TEST INFRASTRUCTURE ONLY -- stand-in for biopython (used by the reference's accuracy benchmark only)."""
SeqIO = None
