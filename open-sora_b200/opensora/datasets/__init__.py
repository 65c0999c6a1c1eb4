"""Only the resolution / aspect-ratio arithmetic the sampler needs (`aspect.py`); datasets, buckets and dataloaders are
the trainer's and out of scope."""
