"""Host-side mirror of the reference's Python module API for the denoising hot path."""
