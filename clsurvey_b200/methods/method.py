"""The Method plugin surface of the reference (src/methods/method.py:81-111, 281-328, 663-757, 994-1087), re-pointed
at the engine for the hot-path methods: EWC, MAS, SI, GEM and the Finetune baseline.

The framework (src/framework/{main,framework_train,lr_grid_train}.py) only ever touches the members kept here:
class attributes name / eval_name / category / extra_hyperparams_count / hyperparams [/ static_hyperparams /
start_scratch / wrap_first_task_model / no_framework / grid_chkpt] and the methods grid_train / train / get_output /
inference_eval [/ poststep].  `parse(name)` returns the same objects, so `framework.main.main(method=parse('EWC'))`
(main.py:77,89-92 accepts injected objects) runs the reference's task loop on this engine.
IMM is provided on the same kernels; LwF, EBLL, PackNet, HAT, iCaRL, the replay baselines and Joint are not.
"""
import copy
import os
from abc import ABC, abstractmethod
from collections import OrderedDict
from enum import Enum, auto

import torch

from ..engine import get_engine
from .EWC import main_EWC as trainEWC
from .Finetune import main_SGD as trainFT
from .IMM import main_L2transfer as trainIMM
from .IMM import merge as mergeIMM
from .MAS import main_MAS as trainMAS
from .rehearsal import main_rehearsal as trainRehearsal
from .SI import main_SI as trainSI


def parse(method_name):
    """method.py:35-78 for the hot-path methods."""
    for cls in (EWC, MAS, SI, GEM, Finetune):
        if method_name == cls.name:
            return cls()
    if IMM.name in method_name:                                       # modeIMM, meanIMM (method.py:73-76)
        return IMM(method_name.replace('_', '').replace(IMM.name, '').strip())
    raise NotImplementedError("Method not on the B200 hot path: %s" % method_name)


class Method(ABC):
    @property
    @abstractmethod
    def name(self): pass

    @property
    @abstractmethod
    def eval_name(self): pass

    @property
    @abstractmethod
    def category(self): pass

    @property
    @abstractmethod
    def extra_hyperparams_count(self): pass

    @property
    @abstractmethod
    def hyperparams(self): pass

    @abstractmethod
    def get_output(self, images, args): pass

    @staticmethod
    @abstractmethod
    def inference_eval(args, manager): pass


class Category(Enum):
    MODEL_BASED = auto()
    DATA_BASED = auto()
    MASK_BASED = auto()
    BASELINE = auto()
    REHEARSAL_BASED = auto()

    def __eq__(self, other):
        return self.name == other.name and self.value == other.value

    __hash__ = Enum.__hash__


def get_output_def(model, heads, images, current_head_idx, final_layer_idx):
    """method.py:230-235: swap in the task's head, eval mode, forward (through the engine)."""
    model.classifier._modules[final_layer_idx] = heads[current_head_idx]
    model.eval()
    eng = get_engine(model, tuple(images.shape[1:]), images.shape[0])
    return eng.forward(images if images.is_cuda else images.to(eng.device), train=False)


def set_hyperparams(method, hyperparams, static_params=False):
    """method.py:238-274: 'a,b;c,d' strings -> (static_)hyperparams OrderedDict."""
    assert isinstance(hyperparams, str)
    leave_default = lambda x: x == 'def' or x == ''
    vals = []
    split_lists = [x.strip() for x in hyperparams.split(';') if len(x) > 0]
    for split_list in split_lists:
        sp = [float(x) for x in split_list.split(',') if not leave_default(x)]
        sp = sp[0] if len(sp) == 1 else sp
        if len(split_lists) == 1:
            vals = sp
        else:
            vals.append(sp)
    if not isinstance(vals, list):
        vals = [vals]
    if static_params:
        if not hasattr(method, 'static_hyperparams'):
            return
        target = method.static_hyperparams
    else:
        target = method.hyperparams
    for i, (key, _) in enumerate(list(target.items())):
        if i < len(vals) and not leave_default(vals[i]):
            target[key] = vals[i]
    method.init_hyperparams = copy.deepcopy(target)


class Finetune(Method):
    """method.py:994-1087.  Deviation: declares no_framework=True like the other baselines, because the
    reference class has no train() and would die in phase 2."""
    name = "finetuning"
    eval_name = name
    category = Category.BASELINE
    extra_hyperparams_count = 0
    hyperparams = {}
    grid_chkpt = True
    start_scratch = True
    no_framework = True

    def get_output(self, images, args):
        return get_output_def(args.model, args.heads, images, args.current_head_idx, args.final_layer_idx)

    @staticmethod
    def grid_train(args, manager, lr):
        dataset_path = manager.current_task_dataset_path
        if not isinstance(dataset_path, list):
            dataset_path = [dataset_path]
        loaders, sizes, classes = Finetune.compose_dataset(dataset_path, args.batch_size)
        return trainFT.fine_tune_SGD(loaders, sizes, classes, model_path=manager.previous_task_model_path,
                                     exp_dir=manager.gridsearch_exp_dir, num_epochs=args.num_epochs, lr=lr,
                                     weight_decay=args.weight_decay, enable_resume=True, save_models_mode=True,
                                     replace_last_classifier_layer=True, freq=args.saving_freq)

    @staticmethod
    def grid_poststep(args, manager):
        manager.previous_task_model_path = os.path.join(manager.best_exp_grid_node_dirname, 'best_model.pth.tar')

    @staticmethod
    def compose_dataset(dataset_path, batch_size):
        """method.py:1043-1063 for a single dataset per task (Joint's multi-dataset concat is out of scope)."""
        assert len(dataset_path) == 1, "dataset concatenation (Joint) is outside the hot path"
        d = dataset_path[0]
        wrapper = torch.load(d, weights_only=False) if isinstance(d, str) else d
        from . import common
        loaders = common.make_loaders(wrapper, batch_size, shuffle=True)      # device-resident task cache (8f-2)
        sizes = {x: len(wrapper[x]) for x in ['train', 'val']}
        classes = {x: [wrapper[x].classes] for x in ['train', 'val']}
        return loaders, sizes, classes

    @staticmethod
    def inference_eval(args, manager):
        from ..framework import inference
        return inference.inference_eval_default(args, manager)


class EWC(Method):
    """method.py:663-692."""
    name = "EWC"
    eval_name = name
    category = Category.MODEL_BASED
    extra_hyperparams_count = 1
    hyperparams = OrderedDict({'lambda': 400})

    @staticmethod
    def grid_train(args, manager, lr):
        return Finetune.grid_train(args, manager, lr)

    def train(self, args, manager, hyperparams):
        return trainEWC.fine_tune_EWC_acuumelation(
            dataset_path=manager.current_task_dataset_path, previous_task_model_path=manager.previous_task_model_path,
            exp_dir=manager.heuristic_exp_dir, data_dir=args.data_dir, reg_sets=manager.reg_sets,
            reg_lambda=hyperparams['lambda'], batch_size=args.batch_size, num_epochs=args.num_epochs, lr=args.lr,
            weight_decay=args.weight_decay, saving_freq=args.saving_freq)

    def get_output(self, images, args):
        return get_output_def(args.model, args.heads, images, args.current_head_idx, args.final_layer_idx)

    @staticmethod
    def inference_eval(args, manager):
        return Finetune.inference_eval(args, manager)


class SI(Method):
    """method.py:695-723."""
    name = "SI"
    eval_name = name
    category = Category.MODEL_BASED
    extra_hyperparams_count = 1
    hyperparams = OrderedDict({'lambda': 400})

    @staticmethod
    def grid_train(args, manager, lr):
        return Finetune.grid_train(args, manager, lr)

    def train(self, args, manager, hyperparams):
        return trainSI.fine_tune_elastic(
            dataset_path=manager.current_task_dataset_path, num_epochs=args.num_epochs,
            exp_dir=manager.heuristic_exp_dir, model_path=manager.previous_task_model_path,
            reg_lambda=hyperparams['lambda'], batch_size=args.batch_size, lr=args.lr, init_freeze=0,
            weight_decay=args.weight_decay, saving_freq=args.saving_freq)

    def get_output(self, images, args):
        return get_output_def(args.model, args.heads, images, args.current_head_idx, args.final_layer_idx)

    @staticmethod
    def inference_eval(args, manager):
        return Finetune.inference_eval(args, manager)


class MAS(Method):
    """method.py:726-757."""
    name = "MAS"
    eval_name = name
    category = Category.MODEL_BASED
    extra_hyperparams_count = 1
    hyperparams = OrderedDict({'lambda': 3})

    @staticmethod
    def grid_train(args, manager, lr):
        return Finetune.grid_train(args, manager, lr)

    def train(self, args, manager, hyperparams):
        return trainMAS.fine_tune_objective_based_acuumelation(
            dataset_path=manager.current_task_dataset_path, previous_task_model_path=manager.previous_task_model_path,
            init_model_path=args.init_model_path, exp_dir=manager.heuristic_exp_dir, data_dir=args.data_dir,
            reg_sets=manager.reg_sets, reg_lambda=hyperparams['lambda'], batch_size=args.batch_size,
            weight_decay=args.weight_decay, num_epochs=args.num_epochs, lr=args.lr, norm='L2', b1=False,
            saving_freq=args.saving_freq)

    def get_output(self, images, args):
        return get_output_def(args.model, args.heads, images, args.current_head_idx, args.final_layer_idx)

    @staticmethod
    def inference_eval(args, manager):
        return Finetune.inference_eval(args, manager)


class IMM(Method):
    """method.py:760-819: L2-transfer training (penalised step with omega = 1), mean / mode merge before
    evaluation."""
    name = "IMM"
    eval_name = name
    modes = ['mean', 'mode']
    category = Category.MODEL_BASED
    extra_hyperparams_count = 1
    hyperparams = OrderedDict({'lambda': 0.01})
    grid_chkpt = True
    no_framework = True

    def __init__(self, mode='mode'):
        if mode not in self.modes:
            raise Exception("NO EXISTING IMM MODE: '{}'".format(mode))
        self.mode = mode
        self.eval_name = self.name + "_" + self.mode

    def set_mode(self, mode):
        if mode not in self.modes:
            raise Exception("TRY TO SET NON EXISTING IMM MODE: ", mode)
        self.mode = mode
        self.eval_name = self.name + "_" + self.mode

    def grid_train(self, args, manager, lr):
        return trainIMM.fine_tune_l2transfer(dataset_path=manager.current_task_dataset_path,
                                             model_path=manager.previous_task_model_path,
                                             exp_dir=manager.gridsearch_exp_dir, reg_lambda=self.hyperparams['lambda'],
                                             batch_size=args.batch_size, num_epochs=args.num_epochs, lr=lr,
                                             weight_decay=args.weight_decay, saving_freq=args.saving_freq)

    def get_output(self, images, args):
        return get_output_def(args.model, args.heads, images, args.current_head_idx, args.final_layer_idx)

    @staticmethod
    def grid_poststep(args, manager):
        manager.previous_task_model_path = os.path.join(manager.best_exp_grid_node_dirname, 'best_model.pth.tar')

    def eval_model_preprocessing(self, args):
        """Merging step before evaluation (method.py:808-813)."""
        return mergeIMM.preprocess_merge_IMM(self, args.models_path, args.datasets_path, args.batch_size, overwrite=True)

    @staticmethod
    def inference_eval(args, manager):
        return Finetune.inference_eval(args, manager)


class GEM(Method):
    """method.py:281-328."""
    name = "GEM"
    eval_name = name
    category = Category.REHEARSAL_BASED
    extra_hyperparams_count = 1
    hyperparams = OrderedDict({'margin': 1})
    static_hyperparams = OrderedDict({'mem_per_task': 1024})
    wrap_first_task_model = True

    def train(self, args, manager, hyperparams):
        return _rehearsal_accespoint(args, manager, hyperparams['margin'], self.static_hyperparams['mem_per_task'], 'gem')

    def get_output(self, images, args):
        o1, o2 = args.model.compute_offsets(args.current_head_idx, args.model.cum_nc_per_task)
        return args.model(images, args.current_head_idx)[:, o1:o2]

    def poststep(self, args, manager):
        if args.task_counter > 1:
            return
        save_path = manager.best_model_path
        if not os.path.exists(save_path):
            _rehearsal_accespoint(args, manager, self.hyperparams['margin'], self.static_hyperparams['mem_per_task'],
                                  'gem', save_path, manager.previous_task_model_path,
                                  postprocess=args.task_counter == 1)
        manager.best_model_path = save_path

    def grid_train(self, args, manager, lr):
        args.lr = lr
        return _rehearsal_accespoint(args, manager, 0, self.static_hyperparams['mem_per_task'], 'gem',
                                     save_path=manager.gridsearch_exp_dir, finetune=True)

    @staticmethod
    def inference_eval(args, manager):
        return Finetune.inference_eval(args, manager)


def _rehearsal_accespoint(args, manager, memory_strength, mem_per_task, method_arg, save_path=None,
                          prev_model_path=None, finetune=False, postprocess=False):
    """method.py:383-413."""
    nc_per_task = list(manager.dataset.nc_per_task) if hasattr(manager.dataset, "nc_per_task") \
        else [manager.dataset.classes_per_task] * manager.dataset.task_count
    manager.overwrite_args = {
        'weight_decay': args.weight_decay, 'task_name': args.task_name, 'task_count': args.task_counter,
        'prev_model_path': manager.previous_task_model_path if prev_model_path is None else prev_model_path,
        'save_path': manager.heuristic_exp_dir if save_path is None else save_path, 'n_outputs': sum(nc_per_task),
        'method': method_arg, 'n_memories': int(mem_per_task), 'n_epochs': args.num_epochs,
        'memory_strength': memory_strength, 'cuda': True, 'dataset_path': manager.current_task_dataset_path,
        'n_tasks': manager.dataset.task_count, 'batch_size': args.batch_size, 'lr': args.lr, 'finetune': finetune,
        'is_scratch_model': args.task_counter == 1, 'postprocess': postprocess,
    }
    return trainRehearsal.main(manager.overwrite_args, nc_per_task)
