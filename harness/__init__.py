"""train / predict harness for the reference's models on the B200 operators -- one process
per GPU (DistributedDataParallel + nn.SyncBatchNorm over NCCL) instead of the reference's single-process
nn.DataParallel (train.py:73), with the reference's command-line flags, loss weights, learning-rate
schedule and checkpoint format.  Not on the hot path; the operators come from `libs/` (ganet_b200)."""
