"""Import stub (rendering is out of scope). ORACLE ONLY."""
