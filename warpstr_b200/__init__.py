"""warpstr_b200 -- B200-native (sm_100a) replacement of WarpSTR's caller hot path."""
__version__ = '0.1.0'
