"""Pix2PixTrainer — mirrors the reference's trainers/pix2pix_trainer.py:11-119 (G step, D step, LR decay,
save) on top of michigan_b200.Pix2PixModel, one process per GPU."""
from .networks.sync_batchnorm import DataParallelWithCallback
from .pix2pix_model import Pix2PixModel


class Pix2PixTrainer:
    def __init__(self, opt):
        self.opt = opt
        self.pix2pix_model = DataParallelWithCallback(Pix2PixModel(opt), device_ids=opt.gpu_ids)
        self.pix2pix_model_on_one_gpu = self.pix2pix_model.module
        self.generated = None
        if opt.isTrain:
            self.optimizer_G, self.optimizer_D = self.pix2pix_model_on_one_gpu.create_optimizers(opt)
            # replaces DataParallel's reduce-add of replica gradients onto GPU 0 (pix2pix_trainer.py:42-43)
            self.pix2pix_model.attach_optimizer(self.optimizer_G)
            self.pix2pix_model.attach_optimizer(self.optimizer_D)
            self.old_lr = opt.lr
        self.g_losses, self.d_losses = {}, {}

    def run_generator_one_step(self, data):
        self.optimizer_G.zero_grad()
        g_losses, generated = self.pix2pix_model(data, mode="generator")
        g_loss = sum(g_losses.values()).mean()
        g_loss.backward()
        self.optimizer_G.step()
        self.g_losses = g_losses
        self.generated = generated

    def run_discriminator_one_step(self, data):
        self.optimizer_D.zero_grad()
        d_losses = self.pix2pix_model(data, mode="discriminator")
        d_loss = sum(d_losses.values()).mean()
        d_loss.backward()
        self.optimizer_D.step()
        self.d_losses = d_losses

    def get_latest_losses(self):
        return {**self.g_losses, **self.d_losses}

    def get_latest_generated(self):
        return self.generated

    def save(self, epoch):
        self.pix2pix_model_on_one_gpu.save(epoch)

    def update_learning_rate(self, epoch):
        if epoch > self.opt.niter:
            new_lr = self.old_lr - self.opt.lr / self.opt.niter_decay
        else:
            new_lr = self.old_lr
        if new_lr != self.old_lr:
            new_lr_G, new_lr_D = (new_lr, new_lr) if self.opt.no_TTUR else (new_lr / 2, new_lr * 2)
            for g in self.optimizer_D.param_groups:
                g["lr"] = new_lr_D
            for g in self.optimizer_G.param_groups:
                g["lr"] = new_lr_G
            print("update learning rate: %f -> %f" % (self.old_lr, new_lr))
            self.old_lr = new_lr
