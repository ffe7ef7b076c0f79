"""Utilities: neighbour lists, units, synthetic benchmark systems."""
