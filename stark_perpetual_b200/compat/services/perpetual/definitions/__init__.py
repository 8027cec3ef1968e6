from stark_perpetual_b200.compat import extend_package_path

extend_package_path(__path__, __name__)
