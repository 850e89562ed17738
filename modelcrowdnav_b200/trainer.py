"""Value-network trainer: mirror of crowd_nav/utils/trainer.py:19-82 (MSE, SGD momentum 0.9, batch 100).

PyTorch autograd does the training (the north star keeps torch "for tensor handoff and training").  Batches are
drawn directly from the device-resident replay tensors instead of a DataLoader over Python tuples; with
`torch.distributed` initialised the gradients are all-reduced as ONE flat bucket (96,502 fp32 = 386 kB, latency
bound) and averaged -- the data-parallel equivalent of the reference's single-process SGD step.
"""
import logging

import torch
import torch.nn as nn
import torch.optim as optim


class Trainer(object):
    def __init__(self, model, memory, device, batch_size, dist_group=None, policy=None):
        self.model = model
        self.device = device
        self.criterion = nn.MSELoss().to(device)
        self.memory = memory
        self.data_loader = None          # kept for interface parity; batches come from the replay tensors
        self.batch_size = batch_size
        self.optimizer = None
        self.dist_group = dist_group
        self.policy = policy             # SARL façade whose GPU lookahead weights are refreshed after training
        self._flat = None

    def set_learning_rate(self, learning_rate):
        logging.info("Current learning rate: %f", learning_rate)
        self.optimizer = optim.SGD(self.model.parameters(), lr=learning_rate, momentum=0.9)

    # -- gradient all-reduce over NVLink (one flat bucket) ----------------------------------------------
    def _sync_gradients(self):
        if self.dist_group is None:
            return
        import torch.distributed as dist
        params = [p for p in self.model.parameters() if p.grad is not None]
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.dist_group)
        flat /= dist.get_world_size(self.dist_group)
        off = 0
        for p in params:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n

    def broadcast_weights(self, src=0):
        """Make every rank start from rank `src`'s weights (after load_state_dict / random init)."""
        if self.dist_group is None:
            return
        import torch.distributed as dist
        for p in self.model.parameters():
            dist.broadcast(p.data, src=src, group=self.dist_group)

    def _step(self, idx):
        inputs = self.memory.states[idx].to(self.device)
        values = self.memory.values[idx].to(self.device)
        self.optimizer.zero_grad()
        outputs = self.model(inputs)
        loss = self.criterion(outputs, values)
        loss.backward()
        self._sync_gradients()
        self.optimizer.step()
        return loss.data.item()

    def _after(self):
        if self.policy is not None:
            self.policy.sync_weights()

    def optimize_epoch(self, num_epochs):
        """trainer.py:36-59: num_epochs passes over a shuffled memory."""
        if self.optimizer is None:
            raise ValueError("Learning rate is not set!")
        n = len(self.memory)
        average_epoch_loss = 0
        for _ in range(num_epochs):
            epoch_loss = 0
            perm = torch.randperm(n, device=self.memory.states.device)
            for s in range(0, n, self.batch_size):
                epoch_loss += self._step(perm[s:s + self.batch_size])
            average_epoch_loss = epoch_loss / n
        self._after()
        return average_epoch_loss

    def optimize_batch(self, num_batches):
        """trainer.py:61-82: each batch is the head of a fresh shuffle (next(iter(DataLoader(shuffle=True))))."""
        if self.optimizer is None:
            raise ValueError("Learning rate is not set!")
        n = len(self.memory)
        losses = 0
        for _ in range(num_batches):
            idx = torch.randperm(n, device=self.memory.states.device)[:self.batch_size]
            losses += self._step(idx)
        average_loss = losses / num_batches
        logging.debug("Average loss : %.2E", average_loss)
        self._after()
        return average_loss
