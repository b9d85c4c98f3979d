"""Host-side mirror of the reference's `models/` package for the coalition-masked hot path."""
