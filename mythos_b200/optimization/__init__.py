"""DiffTRe reweighting objective over the fused energy / parameter-gradient pass."""
