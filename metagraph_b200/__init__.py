"""metagraph_b200: B200-native replacement for the `metagraph align` hot path
(DBGAligner seed-and-extend over a BOSS-encoded succinct de Bruijn graph)."""
from .config import DBGAlignerConfig, cli_defaults, struct_defaults  # noqa: F401
