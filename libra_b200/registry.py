"""Model registry with the reference's interface (libra/common/registry.py:9-19, 54-75, the model half): train.py resolves
`registry.get_model_class(model_config.arch).from_config(model_config)` (train.py:28-30) and the wrapper registers itself as
"libra_train_wrapper" (libra/models/libra/modeling_libra.py:1292)."""


class Registry:
    mapping = {"model_name_mapping": {}, "processor_name_mapping": {}, "state": {}, "paths": {}}

    @classmethod
    def register_model(cls, name):
        def wrap(model_cls):
            if name in cls.mapping["model_name_mapping"]:
                raise KeyError("Name '{}' already registered for {}.".format(name, cls.mapping["model_name_mapping"][name]))
            cls.mapping["model_name_mapping"][name] = model_cls
            return model_cls
        return wrap

    @classmethod
    def get_model_class(cls, name):
        return cls.mapping["model_name_mapping"].get(name, None)

    @classmethod
    def register_processor(cls, name):             # libra/common/registry.py (processor half): "libra_image", "libra_image_eval"
        def wrap(processor_cls):
            if name in cls.mapping["processor_name_mapping"]:
                raise KeyError("Name '{}' already registered for {}.".format(name, cls.mapping["processor_name_mapping"][name]))
            cls.mapping["processor_name_mapping"][name] = processor_cls
            return processor_cls
        return wrap

    @classmethod
    def get_processor_class(cls, name):
        return cls.mapping["processor_name_mapping"].get(name, None)

    @classmethod
    def list_models(cls):
        return sorted(cls.mapping["model_name_mapping"].keys())


registry = Registry()
