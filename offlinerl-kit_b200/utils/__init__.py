"""Host-side utilities with the reference's names: Logger, StandardScaler, termination functions, exploration noise."""
