"""B200-native hot path of keb721/mc_water_ls_mw (mW energy + lattice-switch MC moves)."""
__all__ = ["decks"]
