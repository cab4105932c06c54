"""ccvm_b200 -- B200-native CCVM dynamics engine behind the reference's solver API."""
__version__ = "0.1.0"
