"""reference MFGP_ver2023May/utils/dict_tools.py:3-9: shallow merge of defaults INTO the caller's dict.  It returns
(and mutates) its arguments - with no user config the shared default dict itself is returned, which is what makes the
reference's `[default_config] * n` lists alias one object; reproduced on purpose."""


def update_dict_with_default(default_dict, update_dict):
    if update_dict is None:
        return default_dict
    update_dict.update({key: value for key, value in default_dict.items() if key not in update_dict})
    return update_dict
