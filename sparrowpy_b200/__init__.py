"""sparrowpy_b200 -- B200-native DirectionalRadiosityFast hot path."""
__version__ = "0.1.0"
