"""Input helpers: TOML model constants, oxDNA topology / trajectory readers."""
