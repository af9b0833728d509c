from . import data_types
