from teochat_b200.eval.eval import eval, load_model, main  # noqa: F401,A004

if __name__ == "__main__":
    main()
