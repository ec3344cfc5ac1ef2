"""Import shims for third-party packages the reference needs but the image lacks (test infra only)."""
