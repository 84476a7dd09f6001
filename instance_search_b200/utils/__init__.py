"""Drop-in mirror of the reference's ``utils/`` entry points on the hot path:
``metrics`` (precision1 / avg_precision / mean_avg_precision) and
``train_siamese`` (label indicators, placement rule, all-pairs similarities,
descriptor evaluation)."""
