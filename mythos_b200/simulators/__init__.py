"""Simulator-side pieces of the hot path: trajectory container, fused Langevin integrator, MD run loop."""
