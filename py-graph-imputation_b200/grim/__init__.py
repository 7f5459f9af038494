"""grim -- B200-native drop-in for the imputation hot path of py-graph-imputation.

Same import path and entry points as the reference package (`from grim import grim`;
`grim.impute(conf_file)`, `grim.graph_freqs(conf_file)`), backed by libgrimb200.so."""
__version__ = "0.1.0"
