"""Job settings with the reference's field names, order and defaults (the reference's config.py: module constant at
line 1, class Config below it), kept as a table so that the drop-in contract can be checked field by field."""

simultaneous_tasks_count = 2        # jobs run side by side by the task executor; the reference advises 1 above 2 levels

# (name, default, meaning) in the reference's positional order
FIELDS = (
    ('content_weight', 1e3, 'weight of the content MSE'),
    ('style_weight', 4e5, 'weight of the mean Gram MSE over the five style layers'),
    ('tv_weight', 1e2, 'weight of the total-variation term'),
    ('optimizer', 'lbfgs', "'lbfgs' or 'adam'"),
    ('model', 'vgg19', "feature network; only 'vgg19'"),
    ('init_method', 'content+noise', "'random', 'content+noise' or 'style'"),
    ('levels_num', 2, 'pyramid levels (4 = 2048x3072-class output)'),
    ('iters_num', 500, 'closures to run'),
    ('noise_factor', 0.95, 'strength of the structured noise blended into the initial image'),
    ('noise_levels', (9, 18, 36, -1, 0), 'spots along the short axis (> 0), spot size in px (< 0) or envelope only (0)'),
    ('noise_levels_central_amplitude', (0.30, 0.20, 0.10, 0.20, 0.20), 'Gaussian envelope at the image centre, per level'),
    ('noise_levels_peripheral_amplitude', (0.20, 0.30, 0.40, 0.10, 0.00), 'envelope at the periphery, per level'),
    ('noise_levels_dispersion', (0.20, 0.30, 0.40, 0.60, 0.30), 'envelope sigma as a fraction of the axis, per level'),
)


class Config:
    """Config(**settings) / Config(content_weight, style_weight, ...): attributes named like the reference's."""

    def __init__(self, *args, **kwargs):
        names = [name for name, _, _ in FIELDS]
        if len(args) > len(names):
            raise TypeError(f'Config() takes at most {len(names)} positional arguments ({len(args)} given)')
        values = {name: default for name, default, _ in FIELDS}
        values.update(zip(names, args))
        for key, value in kwargs.items():
            if key not in values:
                raise TypeError(f"Config() got an unexpected keyword argument '{key}'")
            if key in names[:len(args)]:
                raise TypeError(f"Config() got multiple values for argument '{key}'")
            values[key] = value
        for name in names:
            setattr(self, name, values[name])

    def __repr__(self):
        return 'Config(' + ', '.join(f'{name}={getattr(self, name)!r}' for name, _, _ in FIELDS) + ')'
