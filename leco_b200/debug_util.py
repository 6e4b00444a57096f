"""debug_util.py of the reference (print-only probes called at train_lora.py:103-104), same names, for the engine's
parameter-holder tree and the adapter network."""
import torch


def check_requires_grad(model: torch.nn.Module):
    for name, module in list(model.named_modules())[:5]:
        if len(list(module.parameters())) > 0:
            print(f"Module: {name}")
            for pname, param in list(module.named_parameters())[:2]:
                print(f"    Parameter: {pname}, Requires Grad: {param.requires_grad}")


def check_training_mode(model: torch.nn.Module):
    for name, module in list(model.named_modules())[:5]:
        print(f"Module: {name}, Training Mode: {module.training}")
