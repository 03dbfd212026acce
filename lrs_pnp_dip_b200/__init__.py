"""lrs_pnp_dip_b200 (package init is filled in below)."""
