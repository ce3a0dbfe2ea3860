"""Optional sidecar services (AIM metrics driver)."""
